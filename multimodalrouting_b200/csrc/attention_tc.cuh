// tcgen05 / TMEM / TMA cross-attention forward for the bf16 path (multihead_attention.py:93-143), all six directions in
// one launch -- the long-sequence engine (INSPECT token counts, BASELINE configs[4]: 512 / 128 / 196 tokens), selected
// with MMR_ATTN=tc.  The mma.sync kernels of attention_mma.cuh stay the default for the 16-49-token MIMIC shapes, where a
// 128-row UMMA tile would be mostly padding ("Attention").
//
// CTA = (direction, patient, head PAIR, 128-query block), 9 warps:
//   warp 8, lane 0   TMA producer + tcgen05.mma issuer
//                      S_h  = Q_h K_h^T     M=128 queries x N=128 keys x K=32   (two K=16 steps per head, both operands
//                                            K-major slices of one 128B-swizzled [128 x 64] tile that holds both heads)
//                      O_h  = P_h V_pair    M=128 x N=64 x K=128 keys           (A = P_h, K-major, written by the softmax
//                                            warps; B = the [keys x 64] V tile as it lies in memory = MN-major; the 32
//                                            columns of the other head are computed and ignored)
//   warps 0-3 / 4-7  softmax + epilogue of head 0 / 1: thread = query row (TMEM lane), scores read with tcgen05.ld,
//                    rounded to bf16 like the reference's bmm output, key bias added (finfo(bf16).min for padded keys,
//                    -inf for tile padding), online softmax over 128-key chunks with the running output in registers
// NH = 1 variant (the default; MMR_ATTN_TC_HEADS=2 selects head pairs): one head per CTA (5 warps, 80 KB of shared memory, 256 TMEM columns), so two CTAs
// share an SM and one CTA's TMA / MMA / barrier latencies hide behind the other's softmax; the head pair's Q / K / V tiles
// are then fetched by both CTAs of the pair (L2 hits).
// TMEM: S NH x 128 columns, O NH x 64 columns.  Output and statistics are identical in meaning to amma::attn_fwd_kernel
// (unnormalised bf16 P per chunk, final 1/l scaling, ml = (row max, 1/row sum)), so the mma.sync backward consumes them.
//
// Padding: Q / K / V boxes are fetched through per-direction 3-D tensor maps [patient][token][column] whose token extent is
// the direction's T, so the TMA unit zero-fills everything past a patient's last token: no box runs on into the next patient's
// rows (whose scores get a -inf bias and P = 0, but 0 x NaN is NaN inside the MMA).
// Status (profiles/r2_attention_tc.md): parity-green on a B200 at 48/16/49, 150/70/33 and 512/128/196 tokens; 0.70-0.77x the
// mma.sync forward at 512 keys and 7-9x slower at the MIMIC token counts, with or without K / V prefetch -- hence opt-in.
#pragma once
#include "attention_mma.cuh"
#include "gemm_tc.cuh"

namespace mmr {
namespace atc {

using namespace tc;

constexpr int QB = 128;                 // queries per CTA
constexpr int KC = 128;                 // keys per chunk
constexpr int TILE = 128 * 64 * 2;      // one 128B-swizzled [128 rows x 64 bf16] tile
template <int NH> constexpr int threads() { return NH * 128 + 32; }
// Q, K, V, P[NH heads][2 key blocks], key bias, ctrl, alignment
template <int NH, int ST = 1> constexpr int smem_bytes() { return (1 + 2 * ST + 2 * NH) * TILE + 2 * KC * 4 + 256 + 1024; }

struct Ctrl {
  uint64_t kv_full;      // TMA: (Q +) K + V of the chunk landed
  uint64_t s_full;       // both heads' score tiles are in TMEM
  uint64_t p_full[2];    // head's P tile is in shared memory (4 warps arrive)
  uint64_t o_full[2];    // head's P V product is in TMEM
  uint32_t tmem_base;
  uint32_t pad_;
  uint64_t kv_full2;     // ST = 2: the second K / V stage
};

// every legitimate wait in this kernel is far below a millisecond: trap early instead of spinning for minutes
__device__ __forceinline__ void mbar_wait_short(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
template <int N> __device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int x, int y, int z, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

// Q / K / V boxes come from per-direction [patient][token][column] views whose token extent is the direction's T, so that
// the TMA unit zero-fills every row past a patient's last token.
struct Maps3D { CUtensorMap q[NDIR]; CUtensorMap kv[NDIR]; };
struct Load3D {
  const Maps3D* m;
  __device__ __forceinline__ void prefetch(int d) const {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&m->q[d])) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&m->kv[d])) : "memory");
  }
  __device__ __forceinline__ void load_q(uint32_t dst, int col, int tok, int, int b, int d, uint32_t bar) const {
    tma_load_3d(dst, &m->q[d], col, tok, b, bar);
  }
  __device__ __forceinline__ void load_kv(uint32_t dst, int col, int tok, int, int b, int d, uint32_t bar) const {
    tma_load_3d(dst, &m->kv[d], col, tok, b, bar);
  }
};

// ST = 2: two K / V stages (TMA loads of chunk c+1 issued before chunk c is computed); measured slower with one head per CTA and
// equal with two (profiles/r2_attention_tc.md), so only ST = 1 is instantiated.
template <int NH, class Loader, int ST = 1>
__device__ __forceinline__ void attn_fwd_tc_body(const Loader& ldr, const AttnArgs& a) {
  constexpr int CW = NH * 4;            // control warp
  constexpr int HX = 8 / NH;            // CTAs per (patient, query block)
  constexpr uint32_t TMEM_COLS = NH == 2 ? 512 : 256;
  extern __shared__ uint8_t smem_raw[];
  const int d = blockIdx.z, b = blockIdx.y, hx = blockIdx.x % HX, qblk = blockIdx.x / HX;
  const int hp = NH == 2 ? hx : hx >> 1;          // head pair: 64-column slice of Q / K / V
  const int hsel = hx & 1;                        // NH == 1: which head of the pair this CTA owns
  const int Tq = a.q.T[d], Tk = a.kv.T[d];
  const int q0 = qblk * QB;
  if (q0 >= Tq) return;
  const int nq = min(QB, Tq - q0);
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t sQ = sbase, sK = sQ + TILE, sV = sK + TILE, sP = sQ + (1 + 2 * ST) * TILE;   // ST = 2: second K, V after the first
  float* Ms = reinterpret_cast<float*>(sgen + (1 + 2 * ST + 2 * NH) * TILE);   // [2][KC] additive key bias, double buffered
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(Ms + 2 * KC);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&ctrl->kv_full), 1);
    if (ST == 2) mbar_init(smem_u32(&ctrl->kv_full2), 1);
    mbar_init(smem_u32(&ctrl->s_full), 1);
    for (int h = 0; h < NH; ++h) {
      mbar_init(smem_u32(&ctrl->p_full[h]), 4);
      mbar_init(smem_u32(&ctrl->o_full[h]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CW) tmem_alloc(smem_u32(&ctrl->tmem_base), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;
  const uint32_t tS = tmem, tO = tmem + NH * KC;
  const int qrow0 = a.q.row0[d] + b * Tq + q0;
  const int kvrow0 = a.kv.row0[d] + b * Tk;
  const int nchunks = (Tk + KC - 1) / KC;

  if (warp == CW) {
    if (lane == 0) {
      ldr.prefetch(d);
      constexpr uint32_t idesc_s = make_idesc(KC, 0, 0);          // N = 128 keys, both operands K-major
      constexpr uint32_t idesc_o = make_idesc(64, 0, 1);          // N = 64 value columns, B (= V) MN-major
      const uint32_t kv_full = smem_u32(&ctrl->kv_full);
      auto issue_kv = [&](int c, uint32_t bar, uint32_t dk, uint32_t dv) {
        mbar_expect_tx(bar, (c == 0 ? 3u : 2u) * TILE);
        if (c == 0) ldr.load_q(sQ, hp * 64, q0, qrow0, b, d, bar);
        ldr.load_kv(dk, a.col0 + hp * 64, c * KC, kvrow0 + c * KC, b, d, bar);
        ldr.load_kv(dv, a.col0 + D + hp * 64, c * KC, kvrow0 + c * KC, b, d, bar);
      };
      if (ST == 2) issue_kv(0, kv_full, sK, sV);
      for (int c = 0; c < nchunks; ++c) {
        uint32_t sKc = sK, sVc = sV;
        if (ST == 1) {
          // every MMA that read K, V and P of the previous chunk has retired (the second head's commit covers all)
          if (c > 0) mbar_wait_short(smem_u32(&ctrl->o_full[NH - 1]), (uint32_t)(c - 1) & 1u);
          mbar_expect_tx(kv_full, (c == 0 ? 3u : 2u) * TILE);
          if (c == 0) ldr.load_q(sQ, hp * 64, q0, qrow0, b, d, kv_full);
          ldr.load_kv(sK, a.col0 + hp * 64, c * KC, kvrow0 + c * KC, b, d, kv_full);
          ldr.load_kv(sV, a.col0 + D + hp * 64, c * KC, kvrow0 + c * KC, b, d, kv_full);
          mbar_wait_short(kv_full, (uint32_t)c & 1u);
        } else {
          const uint32_t st = (uint32_t)c & 1u;
          if (c + 1 < nchunks) {       // stage of chunk c+1 = stage of chunk c-1: free once that chunk's MMAs have retired
            if (c > 0) mbar_wait_short(smem_u32(&ctrl->o_full[NH - 1]), (uint32_t)(c - 1) & 1u);
            const uint32_t nst = st ^ 1u;
            issue_kv(c + 1, nst ? smem_u32(&ctrl->kv_full2) : kv_full, sK + nst * 2 * TILE, sV + nst * 2 * TILE);
          }
          sKc = sK + st * 2 * TILE; sVc = sV + st * 2 * TILE;
          mbar_wait_short(st ? smem_u32(&ctrl->kv_full2) : kv_full, (uint32_t)(c >> 1) & 1u);
        }
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc(sQ, 16, 1024), kdesc = make_smem_desc(sKc, 16, 1024);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int hc = NH == 2 ? h : hsel;   // head hc = columns [32hc, 32hc+32) of the tile: K=16 steps 2hc, 2hc+1 (+32 B each)
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_bf16(tS + h * KC, qdesc + 2 * (2 * hc + k), kdesc + 2 * (2 * hc + k), idesc_s, k != 0);
        }
        umma_commit(smem_u32(&ctrl->s_full));
        const uint64_t vdesc = make_smem_desc(sVc, 8192, 1024);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          mbar_wait_short(smem_u32(&ctrl->p_full[h]), (uint32_t)c & 1u);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < KC / 16; ++j) {   // 16 keys per step: +32 B inside P's swizzle row, +16 rows (2048 B) of V
            const uint64_t pdesc = make_smem_desc(sP + (2 * h + (j >> 2)) * TILE, 16, 1024) + 2 * (j & 3);
            umma_bf16(tO + h * 64, pdesc, vdesc + 128 * j, idesc_o, j != 0);
          }
          umma_commit(smem_u32(&ctrl->o_full[h]));
        }
      }
    }
  } else {
    const int hl = NH == 2 ? warp >> 2 : 0;        // local head slot: barriers, TMEM accumulators, P tiles
    const int hc = NH == 2 ? hl : hsel;            // which half of the 64-column pair slice
    const int q = warp & 3, r = q * 32 + lane, h = hp * 2 + hc;
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    const uint32_t tS_h = tS + hl * KC + lane_bits;
    const uint32_t tO_h = tO + hl * 64 + hc * 32 + lane_bits;     // this head's 32 columns of its 64-column product
    const uint32_t p_row = sP + (uint32_t)(2 * hl) * TILE + (uint32_t)r * 128;
    const uint32_t sw = (uint32_t)(r & 7);
    const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const int k0 = c * KC;
      float* ms = Ms + (c & 1) * KC;
      if (threadIdx.x < KC) {
        const int idx = k0 + (int)threadIdx.x;
        ms[threadIdx.x] = amma::key_bias(idx < Tk ? (km ? (km[idx] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f);
      }
      softmax_bar<NH * 128>();
      mbar_wait_short(smem_u32(&ctrl->s_full), (uint32_t)c & 1u);
      tc_fence_after();
      float v[32];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < KC / 32; ++j) {
        tmem_ld32(tS_h + j * 32, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, amma::rbf(v[e]) + ms[j * 32 + e]);
      }
      const float mn = fmaxf(m_run, mx);                           // finite: every chunk holds at least one real key
      const float corr = amma::ex2((m_run - mn) * amma::L2E);      // 0 on the first chunk
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < KC / 32; ++j) {
        tmem_ld32(tS_h + j * 32, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          v[e] = amma::ex2((amma::rbf(v[e]) + ms[j * 32 + e] - mn) * amma::L2E);
          sum += v[e];
        }
        // keys [32j, 32j+32) of row r: key block j >> 1, 16-byte units 4 (j & 1) .. +3, XOR-swizzled by r & 7
        const uint32_t blk = p_row + (uint32_t)(j >> 1) * TILE;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t addr = blk + ((((uint32_t)(j & 1) * 4 + u) ^ sw) << 4);
          const uint32_t w0 = pack2_bf16(v[8 * u], v[8 * u + 1]), w1 = pack2_bf16(v[8 * u + 2], v[8 * u + 3]);
          const uint32_t w2 = pack2_bf16(v[8 * u + 4], v[8 * u + 5]), w3 = pack2_bf16(v[8 * u + 6], v[8 * u + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
      }
      l_run = l_run * corr + sum;
      m_run = mn;
      tc_fence_before();
      fence_async_smem();                // generic-proxy writes of P -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&ctrl->p_full[hl]));
      mbar_wait_short(smem_u32(&ctrl->o_full[hl]), (uint32_t)c & 1u);
      tc_fence_after();
      tmem_ld32(tO_h, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = o[i] * corr + v[i];
      tc_fence_before();
    }
    if (r < nq) {
      const float il = 1.0f / l_run;
      bf16* dst = reinterpret_cast<bf16*>(a.o) + (size_t)(qrow0 + r) * D + h * HD;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 w;
        w.x = pack2_bf16(o[8 * u] * il, o[8 * u + 1] * il); w.y = pack2_bf16(o[8 * u + 2] * il, o[8 * u + 3] * il);
        w.z = pack2_bf16(o[8 * u + 4] * il, o[8 * u + 5] * il); w.w = pack2_bf16(o[8 * u + 6] * il, o[8 * u + 7] * il);
        *reinterpret_cast<uint4*>(dst + 8 * u) = w;
      }
      float* p = a.ml + ((size_t)(qrow0 + r) * H + h) * 2;
      p[0] = m_run; p[1] = il;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CW) tmem_dealloc(tmem, TMEM_COLS);
}

template <int NH>
__global__ void __launch_bounds__(threads<NH>(), 3 - NH)
attn_fwd_tc3_kernel(const __grid_constant__ Maps3D maps, AttnArgs a) {
  attn_fwd_tc_body<NH>(Load3D{&maps}, a);
}
// host: one launch for all six directions of a layer
// [patient][token][column] view of one direction's rows inside a [rows, ld] bf16 matrix: box = 64 columns x 128 tokens x 1 patient
static bool make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t tokens, uint64_t patients, uint64_t ld,
                         uint32_t box_tokens) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {cols, tokens, patients};
  cuuint64_t strides[2] = {ld * 2, tokens * ld * 2};
  cuuint32_t box[3] = {64, box_tokens, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NH>
static cudaError_t launch_attn_fwd_tc3_nh(const AttnArgs& a, int B, int maxTq, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc3_kernel<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<NH>());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  Maps3D maps;
  for (int d = 0; d < NDIR; ++d) {
    const bf16* qbase = reinterpret_cast<const bf16*>(a.qb) + (size_t)a.q.row0[d] * D;
    const bf16* kbase = reinterpret_cast<const bf16*>(a.kvbuf) + (size_t)a.kv.row0[d] * a.ldkv;
    if (!make_tmap_3d(&maps.q[d], qbase, (uint64_t)D, (uint64_t)a.q.T[d], (uint64_t)B, (uint64_t)D, QB)) return cudaErrorUnknown;
    if (!make_tmap_3d(&maps.kv[d], kbase, (uint64_t)a.ldkv, (uint64_t)a.kv.T[d], (uint64_t)B, (uint64_t)a.ldkv, KC)) return cudaErrorUnknown;
  }
  dim3 grid((8 / NH) * ((maxTq + QB - 1) / QB), B, NDIR);
  attn_fwd_tc3_kernel<NH><<<grid, threads<NH>(), smem_bytes<NH>(), st>>>(maps, a);
  return cudaGetLastError();
}

// K / V and Q rows are fetched through per-patient 3-D tensor maps (zero fill past a patient's last token), so no box can
// pull in the next patient's rows or pad rows (0 x NaN inside the MMA).  The round-1 2-D-map kernel and its two-stage K / V
// prefetch variant were measured in round 2 (profiles/r2_attention_tc.md: equal or slower) and removed.
static cudaError_t launch_attn_fwd_tc(const AttnArgs& a, int B, int maxTq, cudaStream_t st) {
  const char* eh = getenv("MMR_ATTN_TC_HEADS");     // heads per CTA: 1 (default, two CTAs per SM) or 2
  if (eh && atoi(eh) == 2) return launch_attn_fwd_tc3_nh<2>(a, B, maxTq, st);
  return launch_attn_fwd_tc3_nh<1>(a, B, maxTq, st);
}

}  // namespace atc
}  // namespace mmr
