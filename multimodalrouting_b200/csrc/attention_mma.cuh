// Tensor-core (mma.sync m16n8k16, bf16 -> fp32) per-patient cross attention for the bf16 path,
// forward and backward, all six directions in one launch (multihead_attention.py:93-143).
//
// One CTA handles one (direction, patient, head-group of 4, 64-row chunk): the 128-column slice of
// the Q / K / V (/ dO) rows is staged in shared memory with cp.async (coalesced 256 B row segments),
// fragments come from ldmatrix, and the [rows x keys] score tile lives in registers.  Warp w owns
// head (w & 3) of the group and the 16-row tiles {2*(w>>2), 2*(w>>2)+1} of the chunk.  Sequences
// longer than 64 are walked in 64-row chunks (online softmax in the forward pass).
//
// Masking semantics of the reference: scores are rounded to bf16, padded keys are
// filled with finfo(bf16).min (NOT -inf), softmax is fp32; a patient whose keys are all padded
// attends uniformly.  Tile-padding keys (>= Tk) are excluded outright.
#pragma once
#include "attention.cuh"

namespace mmr {
namespace amma {

// HG = heads per CTA (2 or 4): 2*HG warps, HG*32-column row slices (64*HG bytes), smem row stride HG*32+8
// elements (the +8 keeps ldmatrix conflict-free).  HG=2 gives 128-thread CTAs with ~30-40 KB of shared memory,
// so 4 CTAs per SM overlap each other's load / compute / store phases.
template <int HG> struct Cfg {
  static constexpr int THREADS = HG * 64;
  static constexpr int LDS = HG * 32 + 8;
  static constexpr int NHG = 8 / HG;       // head groups per patient
  static constexpr int CPR = HG * 4;       // 16-byte chunks per staged row
  static constexpr int COLS = HG * 32;
};
constexpr int RC = 64;             // rows (queries or keys) per chunk
constexpr float NEG_BF16 = -3.3895313892515355e38f;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// The reference's bmm emits bf16 scores before its fp32 softmax; the kernels reproduce that rounding point
// (dropping it saves two instructions per score but only 4 % of the attention time, measured).
constexpr bool ROUND_SCORES = true;
__device__ __forceinline__ float rbf(float v) { return ROUND_SCORES ? __bfloat162float(__float2bfloat16_rn(v)) : v; }

// Stage `nvalid` rows x 128 columns (src row stride ld elements) into dst[rows][LDS]; rows
// [nvalid, nrows) are zero-filled so that padded B-operand rows never inject NaN/Inf.
template <int HG>
__device__ __forceinline__ void stage(bf16* dst, const bf16* src, size_t ld, int nvalid, int nrows) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, CPR = Cfg<HG>::CPR;
  for (int idx = threadIdx.x; idx < nrows * CPR; idx += THREADS) {
    const int r = idx / CPR, c = (idx % CPR) * 8;
    bf16* d = dst + r * LDS + c;
    if (r < nvalid) cp_async16(smem_addr(d), src + (size_t)r * ld + c);
    else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
  }
}

// Rows [nv, pad256(nv)) of a packed q-space tensor must read as zeros (the weight-gradient GEMMs reduce over them).  The
// CTAs of patient 0 clear them for their own column slice of the attention output / dQ, which replaces a separate
// zero_pad_rows launch after every attention launch (8 per step).  No-op for dense row spaces without tile padding.
template <int HG>
__device__ __forceinline__ void zero_pad_slice(const Segs& q, int d, bf16* slice, int ld) {
  constexpr int THREADS = Cfg<HG>::THREADS, CPR = Cfg<HG>::CPR;
  const int r0 = seg_rows(q, d), n = seg_rows_z(q, d) - r0;
  for (int idx = threadIdx.x; idx < n * CPR; idx += THREADS) {
    const int r = r0 + idx / CPR, c = (idx % CPR) * 8;
    *reinterpret_cast<uint4*>(slice + ((size_t)q.row0[d] + r) * ld + c) = make_uint4(0, 0, 0, 0);
  }
}

// A fragment (16 rows x 16 k) of a row-major smem tile at (row0, col0)
template <int LDS>
__device__ __forceinline__ void frag_a(const bf16* s, int row0, int col0, int lane, uint32_t (&r)[4]) {
  ldsm_x4(smem_addr(s + (row0 + (lane & 15)) * LDS + col0 + (lane >> 4) * 8), r);
}
// B fragments for C[m][n] += A[m][k] * X[n][k] (X row-major, n = row, k = col): 8 rows n0.., k = col0..col0+31
//   r[0],r[1] = (b0,b1) of k-step 0, r[2],r[3] = (b0,b1) of k-step 1
template <int LDS>
__device__ __forceinline__ void frag_b_nk(const bf16* s, int n0, int col0, int lane, uint32_t (&r)[4]) {
  ldsm_x4(smem_addr(s + (n0 + (lane & 7)) * LDS + col0 + (lane >> 3) * 8), r);
}
// B fragments for C[m][n] += A[m][k] * X[k][n] (X row-major, k = row, n = col): 16 rows k0.., cols n0..n0+15
//   r[0],r[1] = (b0,b1) of n-tile n0, r[2],r[3] = (b0,b1) of n-tile n0+8
template <int LDS>
__device__ __forceinline__ void frag_b_kn(const bf16* s, int k0, int n0, int lane, uint32_t (&r)[4]) {
  ldsm_x4_t(smem_addr(s + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + n0 + (lane >> 4) * 8), r);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float L2E = 1.4426950408889634f;
// additive key bias: 0 for a kept key, finfo(bf16).min for a padded key (a bf16 score plus it IS finfo.min in
// fp32, i.e. masked_fill), -inf for tile padding beyond Tk
__device__ __forceinline__ float key_bias(float mk) { return mk > 0.f ? 0.f : (mk == 0.f ? NEG_BF16 : -INFINITY); }

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// copy `nrows` x 128 columns from smem tile to global (row stride ld), coalesced 16-byte stores
template <int HG>
__device__ __forceinline__ void unstage(bf16* dst, size_t ld, const bf16* src, int nrows) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, CPR = Cfg<HG>::CPR;
  for (int idx = threadIdx.x; idx < nrows * CPR; idx += THREADS) {
    const int r = idx / CPR, c = (idx % CPR) * 8;
    *reinterpret_cast<uint4*>(dst + (size_t)r * ld + c) = *reinterpret_cast<const uint4*>(src + r * LDS + c);
  }
}

// write a 16x8 fp32 C fragment tile as bf16 into smem at (row0, col0)
template <int LDS>
__device__ __forceinline__ void put_c(bf16* s, int row0, int col0, int lane, const float (&c)[4], float s0 = 1.f, float s1 = 1.f) {
  const int g = lane >> 2, t = lane & 3;
  *reinterpret_cast<uint32_t*>(s + (row0 + g) * LDS + col0 + 2 * t) = pack_bf16(c[0] * s0, c[1] * s0);
  *reinterpret_cast<uint32_t*>(s + (row0 + g + 8) * LDS + col0 + 2 * t) = pack_bf16(c[2] * s1, c[3] * s1);
}

template <int HG> constexpr int fwd_smem() { return 3 * RC * Cfg<HG>::LDS * 2 + RC * 4; }
template <int HG> constexpr int bwd_smem() { return 4 * RC * Cfg<HG>::LDS * 2 + RC * 4 + RC * HG * 3 * 4; }

// grid: (NHG * ceil(maxTq/64), B, 6)
template <int HG>
__global__ void __launch_bounds__(Cfg<HG>::THREADS, 512 / Cfg<HG>::THREADS) attn_fwd_kernel(AttnArgs a) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, NHG = Cfg<HG>::NHG, COLS = Cfg<HG>::COLS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + RC * LDS;
  bf16* Vs = Ks + RC * LDS;
  float* Ms = reinterpret_cast<float*>(Vs + RC * LDS);   // [64] additive key bias
  pdl_trigger();
  pdl_wait();
  const int d = blockIdx.z, b = blockIdx.y, hg = blockIdx.x % NHG, qc = blockIdx.x / NHG;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  const int q0 = qc * RC;
  if (q0 >= Tq) return;
  const int nq = min(RC, Tq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = warp % HG, half = warp / HG, h = hg * HG + hl;
  const int g = lane >> 2, t = lane & 3;
  const size_t qrow0 = (size_t)a.q.row0[d] + (size_t)qs_ + q0;
  const bf16* qsrc = reinterpret_cast<const bf16*>(a.qb) + qrow0 * D + hg * COLS;
  const bf16* kvsrc = reinterpret_cast<const bf16*>(a.kvbuf) + ((size_t)a.kv.row0[d] + (size_t)b * Tk) * a.ldkv + a.col0 + hg * COLS;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
  const int nq16 = (nq + 15) & ~15;
  stage<HG>(Qs, qsrc, D, nq, nq16);
  const bool single = Tk <= RC;

  float o[2][4][4];
  float mrun[2][2], lrun[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    mrun[i][0] = mrun[i][1] = -INFINITY;
    lrun[i][0] = lrun[i][1] = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][n][j] = 0.f;
  }

  for (int k0 = 0; k0 < Tk; k0 += RC) {
    const int nk = min(RC, Tk - k0);
    const int nk16 = (nk + 15) & ~15;
    __syncthreads();   // previous chunk fully consumed
    stage<HG>(Ks, kvsrc + (size_t)k0 * a.ldkv, a.ldkv, nk, nk16);
    stage<HG>(Vs, kvsrc + (size_t)k0 * a.ldkv + D, a.ldkv, nk, nk16);
    if (threadIdx.x < RC)
      Ms[threadIdx.x] = key_bias(threadIdx.x < nk ? (km ? (km[k0 + threadIdx.x] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f);
    cp_async_wait_all();
    __syncthreads();
    const int NT = (nk + 7) >> 3;      // 8-key score tiles that hold at least one real key
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int mt = half * 2 + i;
      if (mt * 16 >= nq) continue;
      uint32_t qa[2][4];
      frag_a<LDS>(Qs, mt * 16, hl * 32, lane, qa[0]);
      frag_a<LDS>(Qs, mt * 16, hl * 32 + 16, lane, qa[1]);
      float s[8][4];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        if (nt < NT) {
          uint32_t kb[4];
          frag_b_nk<LDS>(Ks, nt * 8, hl * 32, lane, kb);
          mma16816(s[nt], qa[0], kb[0], kb[1]);
          mma16816(s[nt], qa[1], kb[2], kb[3]);
          const float2 kbias = *reinterpret_cast<const float2*>(Ms + nt * 8 + 2 * t);
          s[nt][0] = rbf(s[nt][0]) + kbias.x; s[nt][1] = rbf(s[nt][1]) + kbias.y;
          s[nt][2] = rbf(s[nt][2]) + kbias.x; s[nt][3] = rbf(s[nt][3]) + kbias.y;
          mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      const float mn0 = fmaxf(mrun[i][0], mx0), mn1 = fmaxf(mrun[i][1], mx1);
      const float c0 = ex2((mrun[i][0] - mn0) * L2E), c1 = ex2((mrun[i][1] - mn1) * L2E);   // 0 on the first chunk
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        if (nt < NT) {
          s[nt][0] = ex2((s[nt][0] - mn0) * L2E); s[nt][1] = ex2((s[nt][1] - mn0) * L2E);
          s[nt][2] = ex2((s[nt][2] - mn1) * L2E); s[nt][3] = ex2((s[nt][3] - mn1) * L2E);
          sum0 += s[nt][0] + s[nt][1];
          sum1 += s[nt][2] + s[nt][3];
        }
      }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      lrun[i][0] = lrun[i][0] * c0 + sum0;
      lrun[i][1] = lrun[i][1] * c1 + sum1;
      mrun[i][0] = mn0; mrun[i][1] = mn1;
      float ps0 = 1.f, ps1 = 1.f;
      if (single) { ps0 = 1.0f / sum0; ps1 = 1.0f / sum1; }   // exact reference order: normalise, then round to bf16
      else {
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[i][n][0] *= c0; o[i][n][1] *= c0; o[i][n][2] *= c1; o[i][n][3] *= c1; }
      }
      // O += P V   (score tiles >= NT are exactly zero)
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (2 * kk < NT) {
          uint32_t pa[4];
          pa[0] = pack_bf16(s[2 * kk][0] * ps0, s[2 * kk][1] * ps0);
          pa[1] = pack_bf16(s[2 * kk][2] * ps1, s[2 * kk][3] * ps1);
          pa[2] = pack_bf16(s[2 * kk + 1][0] * ps0, s[2 * kk + 1][1] * ps0);
          pa[3] = pack_bf16(s[2 * kk + 1][2] * ps1, s[2 * kk + 1][3] * ps1);
#pragma unroll
          for (int nc = 0; nc < 2; ++nc) {
            uint32_t vb[4];
            frag_b_kn<LDS>(Vs, kk * 16, hl * 32 + nc * 16, lane, vb);
            mma16816(o[i][2 * nc], pa, vb[0], vb[1]);
            mma16816(o[i][2 * nc + 1], pa, vb[2], vb[3]);
          }
        }
      }
    }
  }
  // epilogue: normalise, stage O through the (warp-private) Q slots, write statistics
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nq) continue;
    const float il0 = 1.0f / lrun[i][0], il1 = 1.0f / lrun[i][1];
    const float s0 = single ? 1.f : il0, s1 = single ? 1.f : il1;
#pragma unroll
    for (int n = 0; n < 4; ++n) put_c<LDS>(Qs, mt * 16, hl * 32 + n * 8, lane, o[i][n], s0, s1);
    if (t == 0) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      if (r0 < nq) { float* p = a.ml + ((qrow0 + r0) * H + h) * 2; p[0] = mrun[i][0]; p[1] = il0; }
      if (r1 < nq) { float* p = a.ml + ((qrow0 + r1) * H + h) * 2; p[0] = mrun[i][1]; p[1] = il1; }
    }
  }
  __syncthreads();
  unstage<HG>(reinterpret_cast<bf16*>(a.o) + qrow0 * D + hg * COLS, D, Qs, nq);
}

// ------------------------------------------------------------------------------------------
// Forward specialised for Tk <= 64 (one key chunk: every MIMIC-IV shape).  No running max / sum / output
// state survives an m-tile, so the tiles are walked in a rolled loop and the kernel fits 6 CTAs (24 warps)
// per SM -- the general kernel above is bound by dependent-instruction latency at 16 warps per SM.
// grid: (NHG * ceil(maxTq/64), B, 6)
template <int HG>
__global__ void __launch_bounds__(Cfg<HG>::THREADS, 768 / Cfg<HG>::THREADS) attn_fwd_single_kernel(AttnArgs a) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, NHG = Cfg<HG>::NHG, COLS = Cfg<HG>::COLS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Ks = Qs + RC * LDS;
  bf16* Vs = Ks + RC * LDS;
  float* Ms = reinterpret_cast<float*>(Vs + RC * LDS);   // [64] additive key bias
  const int d = blockIdx.z, b = blockIdx.y, hg = blockIdx.x % NHG, qc = blockIdx.x / NHG;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int nk = a.kv.T[d];
  const int q0 = qc * RC;
  if (b == 0 && qc == 0) zero_pad_slice<HG>(a.q, d, reinterpret_cast<bf16*>(a.o) + hg * COLS, D);
  if (q0 >= Tq) return;
  const int nq = min(RC, Tq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = warp % HG, half = warp / HG, h = hg * HG + hl;
  const int g = lane >> 2, t = lane & 3;
  const size_t qrow0 = (size_t)a.q.row0[d] + (size_t)qs_ + q0;
  const bf16* kvsrc = reinterpret_cast<const bf16*>(a.kvbuf) + ((size_t)a.kv.row0[d] + (size_t)b * nk) * a.ldkv + a.col0 + hg * COLS;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * nk : nullptr;
  const int nq16 = (nq + 15) & ~15, nk16 = (nk + 15) & ~15;
  stage<HG>(Qs, reinterpret_cast<const bf16*>(a.qb) + qrow0 * D + hg * COLS, D, nq, nq16);
  stage<HG>(Ks, kvsrc, a.ldkv, nk, nk16);
  stage<HG>(Vs, kvsrc + D, a.ldkv, nk, nk16);
  if (threadIdx.x < RC)
    Ms[threadIdx.x] = key_bias(threadIdx.x < nk ? (km ? (km[threadIdx.x] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f);
  cp_async_wait_all();
  __syncthreads();
  const int NT = (nk + 7) >> 3;
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nq) break;
    uint32_t qa[2][4];
    frag_a<LDS>(Qs, mt * 16, hl * 32, lane, qa[0]);
    frag_a<LDS>(Qs, mt * 16, hl * 32 + 16, lane, qa[1]);
    float s[8][4];
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      if (nt < NT) {
        uint32_t kb[4];
        frag_b_nk<LDS>(Ks, nt * 8, hl * 32, lane, kb);
        mma16816(s[nt], qa[0], kb[0], kb[1]);
        mma16816(s[nt], qa[1], kb[2], kb[3]);
        const float2 kbias = *reinterpret_cast<const float2*>(Ms + nt * 8 + 2 * t);
        s[nt][0] = rbf(s[nt][0]) + kbias.x; s[nt][1] = rbf(s[nt][1]) + kbias.y;
        s[nt][2] = rbf(s[nt][2]) + kbias.x; s[nt][3] = rbf(s[nt][3]) + kbias.y;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
    }
    mx0 = quad_max(mx0); mx1 = quad_max(mx1);
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < NT) {
        s[nt][0] = ex2((s[nt][0] - mx0) * L2E); s[nt][1] = ex2((s[nt][1] - mx0) * L2E);
        s[nt][2] = ex2((s[nt][2] - mx1) * L2E); s[nt][3] = ex2((s[nt][3] - mx1) * L2E);
        sum0 += s[nt][0] + s[nt][1];
        sum1 += s[nt][2] + s[nt][3];
      }
    }
    sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
    const float il0 = 1.0f / sum0, il1 = 1.0f / sum1;      // reference order: normalise, then round P to bf16
    float o[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (2 * kk < NT) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kk][0] * il0, s[2 * kk][1] * il0);
        pa[1] = pack_bf16(s[2 * kk][2] * il1, s[2 * kk][3] * il1);
        pa[2] = pack_bf16(s[2 * kk + 1][0] * il0, s[2 * kk + 1][1] * il0);
        pa[3] = pack_bf16(s[2 * kk + 1][2] * il1, s[2 * kk + 1][3] * il1);
#pragma unroll
        for (int nc = 0; nc < 2; ++nc) {
          uint32_t vb[4];
          frag_b_kn<LDS>(Vs, kk * 16, hl * 32 + nc * 16, lane, vb);
          mma16816(o[2 * nc], pa, vb[0], vb[1]);
          mma16816(o[2 * nc + 1], pa, vb[2], vb[3]);
        }
      }
    }
    // this warp is the only reader of its (m-tile, head) slot of Qs: reuse it as the output staging tile
    __syncwarp();
#pragma unroll
    for (int n = 0; n < 4; ++n) put_c<LDS>(Qs, mt * 16, hl * 32 + n * 8, lane, o[n]);
    if (t == 0) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      if (r0 < nq) { float* p = a.ml + ((qrow0 + r0) * H + h) * 2; p[0] = mx0; p[1] = il0; }
      if (r1 < nq) { float* p = a.ml + ((qrow0 + r1) * H + h) * 2; p[0] = mx1; p[1] = il1; }
    }
  }
  __syncthreads();
  unstage<HG>(reinterpret_cast<bf16*>(a.o) + qrow0 * D + hg * COLS, D, Qs, nq);
}

// ------------------------------------------------------------------------------------------
// dQ pass.  CTA = (direction, patient, head group, 64-query chunk); loops over key chunks.
//   P = exp(S - m) / l ; dP = dO V^T ; dS = P .* (dP - D) on kept keys ; dQ = dS K
// Also writes D = rowsum(dO .* O) to a.dvec for the dK/dV pass.
template <int HG>
__global__ void __launch_bounds__(Cfg<HG>::THREADS, 512 / Cfg<HG>::THREADS) attn_bwd_dq_kernel(AttnArgs a) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, NHG = Cfg<HG>::NHG, COLS = Cfg<HG>::COLS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Gs = Qs + RC * LDS;      // dO
  bf16* Ks = Gs + RC * LDS;
  bf16* Vs = Ks + RC * LDS;
  float* Ms = reinterpret_cast<float*>(Vs + RC * LDS);
  float* St = Ms + RC;           // [64][HG][3]: m, 1/l, D
  const int d = blockIdx.z, b = blockIdx.y, hg = blockIdx.x % NHG, qc = blockIdx.x / NHG;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  const int q0 = qc * RC;
  if (q0 >= Tq) return;
  const int nq = min(RC, Tq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = warp % HG, half = warp / HG;
  const int g = lane >> 2, t = lane & 3;
  const size_t qrow0 = (size_t)a.q.row0[d] + (size_t)qs_ + q0;
  const bf16* kvsrc = reinterpret_cast<const bf16*>(a.kvbuf) + ((size_t)a.kv.row0[d] + (size_t)b * Tk) * a.ldkv + a.col0 + hg * COLS;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
  const int nq16 = (nq + 15) & ~15;
  stage<HG>(Qs, reinterpret_cast<const bf16*>(a.qb) + qrow0 * D + hg * COLS, D, nq, nq16);
  stage<HG>(Gs, reinterpret_cast<const bf16*>(a.d_o) + qrow0 * D + hg * COLS, D, nq, nq16);
  // row statistics: (row, head) pairs of this CTA; D from global O / dO (64-byte head slices)
  for (int idx = threadIdx.x; idx < nq * HG; idx += THREADS) {
    const int r = idx / HG, hh = idx % HG;
    const size_t row = qrow0 + r;
    const int hglob = hg * HG + hh;
    const uint4* po = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.o) + row * D + hglob * HD);
    const uint4* pg = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.d_o) + row * D + hglob * HD);
    float dv = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 x = po[i], y = pg[i];
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dv = fmaf(__uint_as_float(xs[j] << 16), __uint_as_float(ys[j] << 16), dv);
        dv = fmaf(__uint_as_float(xs[j] & 0xffff0000u), __uint_as_float(ys[j] & 0xffff0000u), dv);
      }
    }
    const float* ml = a.ml + (row * H + hglob) * 2;
    St[idx * 3] = ml[0]; St[idx * 3 + 1] = ml[1]; St[idx * 3 + 2] = dv;
    a.dvec[row * H + hglob] = dv;
  }
  float dq[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) dq[i][n][j] = 0.f;

  for (int k0 = 0; k0 < Tk; k0 += RC) {
    const int nk = min(RC, Tk - k0);
    const int nk16 = (nk + 15) & ~15;
    __syncthreads();
    stage<HG>(Ks, kvsrc + (size_t)k0 * a.ldkv, a.ldkv, nk, nk16);
    stage<HG>(Vs, kvsrc + (size_t)k0 * a.ldkv + D, a.ldkv, nk, nk16);
    if (threadIdx.x < RC) Ms[threadIdx.x] = threadIdx.x < nk ? (km ? (km[k0 + threadIdx.x] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f;
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int mt = half * 2 + i;
      if (mt * 16 >= nq) continue;
      uint32_t qa[2][4], ga[2][4];
      frag_a<LDS>(Qs, mt * 16, hl * 32, lane, qa[0]);
      frag_a<LDS>(Qs, mt * 16, hl * 32 + 16, lane, qa[1]);
      frag_a<LDS>(Gs, mt * 16, hl * 32, lane, ga[0]);
      frag_a<LDS>(Gs, mt * 16, hl * 32 + 16, lane, ga[1]);
      const int r0 = min(mt * 16 + g, nq - 1), r1 = min(mt * 16 + g + 8, nq - 1);
      const float m0 = St[(r0 * HG + hl) * 3], il0 = St[(r0 * HG + hl) * 3 + 1], D0 = St[(r0 * HG + hl) * 3 + 2];
      const float m1 = St[(r1 * HG + hl) * 3], il1 = St[(r1 * HG + hl) * 3 + 1], D1 = St[(r1 * HG + hl) * 3 + 2];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk * 16 >= nk16) continue;
        uint32_t dsa[4];   // dS of this 16-key step as an A fragment
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int nt = 2 * kk + e;
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t kb[4], vb[4];
          frag_b_nk<LDS>(Ks, nt * 8, hl * 32, lane, kb);
          frag_b_nk<LDS>(Vs, nt * 8, hl * 32, lane, vb);
          mma16816(s, qa[0], kb[0], kb[1]);
          mma16816(s, qa[1], kb[2], kb[3]);
          mma16816(dp, ga[0], vb[0], vb[1]);
          mma16816(dp, ga[1], vb[2], vb[3]);
          float ds[4];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float mk = Ms[nt * 8 + 2 * t + j];
            const float p0 = __expf(rbf(s[j]) - m0) * il0, p1 = __expf(rbf(s[2 + j]) - m1) * il1;
            ds[j] = mk > 0.f ? p0 * (dp[j] - D0) : 0.f;
            ds[2 + j] = mk > 0.f ? p1 * (dp[2 + j] - D1) : 0.f;
          }
          dsa[e * 2] = pack_bf16(ds[0], ds[1]);
          dsa[e * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
        for (int nc = 0; nc < 2; ++nc) {
          uint32_t kb[4];
          frag_b_kn<LDS>(Ks, kk * 16, hl * 32 + nc * 16, lane, kb);
          mma16816(dq[i][2 * nc], dsa, kb[0], kb[1]);
          mma16816(dq[i][2 * nc + 1], dsa, kb[2], kb[3]);
        }
      }
    }
  }
  __syncthreads();   // all reads of Qs done before it is reused as the dQ staging tile
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nq) continue;
#pragma unroll
    for (int n = 0; n < 4; ++n) put_c<LDS>(Qs, mt * 16, hl * 32 + n * 8, lane, dq[i][n]);
  }
  __syncthreads();
  unstage<HG>(reinterpret_cast<bf16*>(a.dq) + qrow0 * D + hg * COLS, D, Qs, nq);
}

// ------------------------------------------------------------------------------------------
// dK / dV pass (transposed tiles).  CTA = (direction, patient, head group, 64-key chunk); loops
// over query chunks.  S^T = K Q^T ; dP^T = V dO^T ; dV = P^T dO ; dK = dS^T Q.
// A padded key still receives dV (its probability is only zero when another key is kept) but no dK.
template <int HG>
__global__ void __launch_bounds__(Cfg<HG>::THREADS, 512 / Cfg<HG>::THREADS) attn_bwd_dkv_kernel(AttnArgs a) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, NHG = Cfg<HG>::NHG, COLS = Cfg<HG>::COLS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + RC * LDS;
  bf16* Qs = Vs + RC * LDS;
  bf16* Gs = Qs + RC * LDS;
  float* Ms = reinterpret_cast<float*>(Gs + RC * LDS);
  float* St = Ms + RC;           // [64 queries][HG][3]
  const int d = blockIdx.z, b = blockIdx.y, hg = blockIdx.x % NHG, kc = blockIdx.x / NHG;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  const int k0 = kc * RC;
  if (k0 >= Tk) return;
  const int nk = min(RC, Tk - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = warp % HG, half = warp / HG;
  const int g = lane >> 2, t = lane & 3;
  const size_t krow0 = (size_t)a.kv.row0[d] + (size_t)b * Tk + k0;
  const size_t qbase = (size_t)a.q.row0[d] + (size_t)qs_;
  const bf16* kvsrc = reinterpret_cast<const bf16*>(a.kvbuf) + krow0 * a.ldkv + a.col0 + hg * COLS;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk + k0 : nullptr;
  const int nk16 = (nk + 15) & ~15;
  stage<HG>(Ks, kvsrc, a.ldkv, nk, nk16);
  stage<HG>(Vs, kvsrc + D, a.ldkv, nk, nk16);
  if (threadIdx.x < RC) Ms[threadIdx.x] = threadIdx.x < nk ? (km ? (km[threadIdx.x] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f;

  float dk[2][4][4], dv[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) { dk[i][n][j] = 0.f; dv[i][n][j] = 0.f; }

  for (int q0 = 0; q0 < Tq; q0 += RC) {
    const int nq = min(RC, Tq - q0);
    const int nq16 = (nq + 15) & ~15;
    __syncthreads();
    stage<HG>(Qs, reinterpret_cast<const bf16*>(a.qb) + (qbase + q0) * D + hg * COLS, D, nq, nq16);
    stage<HG>(Gs, reinterpret_cast<const bf16*>(a.d_o) + (qbase + q0) * D + hg * COLS, D, nq, nq16);
    for (int idx = threadIdx.x; idx < RC * HG; idx += THREADS) {
      const int r = idx / HG, hh = idx % HG;
      float m = 0.f, il = 0.f, dd = 0.f;
      if (r < nq) {
        const size_t rr = (qbase + q0 + r) * H + hg * HG + hh;
        m = a.ml[rr * 2]; il = a.ml[rr * 2 + 1]; dd = a.dvec[rr];
      }
      St[idx * 3] = m; St[idx * 3 + 1] = il; St[idx * 3 + 2] = dd;   // il = 0 for tile-padding queries -> P = 0
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int mt = half * 2 + i;
      if (mt * 16 >= nk) continue;
      uint32_t ka[2][4], va[2][4];
      frag_a<LDS>(Ks, mt * 16, hl * 32, lane, ka[0]);
      frag_a<LDS>(Ks, mt * 16, hl * 32 + 16, lane, ka[1]);
      frag_a<LDS>(Vs, mt * 16, hl * 32, lane, va[0]);
      frag_a<LDS>(Vs, mt * 16, hl * 32 + 16, lane, va[1]);
      const float mk0 = Ms[mt * 16 + g], mk1 = Ms[mt * 16 + g + 8];   // key flags of this thread's two rows
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (kk * 16 >= nq16) continue;
        uint32_t pa[4], dsa[4];   // P^T and dS^T of this 16-query step as A fragments
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int nt = 2 * kk + e;
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t qb[4], gb[4];
          frag_b_nk<LDS>(Qs, nt * 8, hl * 32, lane, qb);
          frag_b_nk<LDS>(Gs, nt * 8, hl * 32, lane, gb);
          mma16816(s, ka[0], qb[0], qb[1]);
          mma16816(s, ka[1], qb[2], qb[3]);
          mma16816(dp, va[0], gb[0], gb[1]);
          mma16816(dp, va[1], gb[2], gb[3]);
          float p[4], ds[4];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const float* st = St + ((nt * 8 + 2 * t + j) * HG + hl) * 3;
            const float m = st[0], il = st[1], dd = st[2];
            float s0 = rbf(s[j]), s1 = rbf(s[2 + j]);
            if (mk0 == 0.f) s0 = NEG_BF16;
            if (mk1 == 0.f) s1 = NEG_BF16;
            const float p0 = mk0 < 0.f ? 0.f : __expf(s0 - m) * il;
            const float p1 = mk1 < 0.f ? 0.f : __expf(s1 - m) * il;
            p[j] = p0; p[2 + j] = p1;
            ds[j] = mk0 > 0.f ? p0 * (dp[j] - dd) : 0.f;
            ds[2 + j] = mk1 > 0.f ? p1 * (dp[2 + j] - dd) : 0.f;
          }
          pa[e * 2] = pack_bf16(p[0], p[1]);
          pa[e * 2 + 1] = pack_bf16(p[2], p[3]);
          dsa[e * 2] = pack_bf16(ds[0], ds[1]);
          dsa[e * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
#pragma unroll
        for (int nc = 0; nc < 2; ++nc) {
          uint32_t gb[4], qb[4];
          frag_b_kn<LDS>(Gs, kk * 16, hl * 32 + nc * 16, lane, gb);
          frag_b_kn<LDS>(Qs, kk * 16, hl * 32 + nc * 16, lane, qb);
          mma16816(dv[i][2 * nc], pa, gb[0], gb[1]);
          mma16816(dv[i][2 * nc + 1], pa, gb[2], gb[3]);
          mma16816(dk[i][2 * nc], dsa, qb[0], qb[1]);
          mma16816(dk[i][2 * nc + 1], dsa, qb[2], qb[3]);
        }
      }
    }
  }
  cp_async_wait_all();     // a patient without query rows (packed layout) never entered the loop: K / V may still be landing
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nk) continue;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      put_c<LDS>(Ks, mt * 16, hl * 32 + n * 8, lane, dk[i][n]);
      put_c<LDS>(Vs, mt * 16, hl * 32 + n * 8, lane, dv[i][n]);
    }
  }
  __syncthreads();
  bf16* out = reinterpret_cast<bf16*>(a.dkv) + krow0 * a.ldkv + a.col0 + hg * COLS;
  unstage<HG>(out, a.ldkv, Ks, nk);
  unstage<HG>(out + D, a.ldkv, Vs, nk);
}

// ------------------------------------------------------------------------------------------
// Fused backward for sequences that fit one chunk (Tq <= 64 and Tk <= 64: all MIMIC-IV shapes).
// CTA = (direction, patient, head group): Q, dO, K, V are staged ONCE; pass A (rows = queries) produces dQ
// and the softmax-backward row term D_i = sum_j P_ij dP_ij (== rowsum(dO .* O), so O is never re-read);
// pass B (rows = keys, transposed tiles) produces dK and dV.  grid: (NHG, B, 6)
template <int HG> constexpr int bwd_fused_smem() { return 4 * RC * Cfg<HG>::LDS * 2 + 2 * RC * 4 + RC * HG * 4 * 4; }

// write a 16x8 fp32 C fragment tile as bf16 straight to global memory (row stride ld), rows < nrows only
__device__ __forceinline__ void put_c_global(bf16* dst, size_t ld, int row0, int col0, int lane, const float (&c)[4], int nrows) {
  const int g = lane >> 2, t = lane & 3;
  if (row0 + g < nrows) *reinterpret_cast<uint32_t*>(dst + (size_t)(row0 + g) * ld + col0 + 2 * t) = pack_bf16(c[0], c[1]);
  if (row0 + g + 8 < nrows) *reinterpret_cast<uint32_t*>(dst + (size_t)(row0 + g + 8) * ld + col0 + 2 * t) = pack_bf16(c[2], c[3]);
}

// The m-tile loops are rolled and P is held packed (bf16x2) between the row-sum and the dS product, so the kernel
// fits 5 CTAs (20 warps) per SM: like the forward it is bound by dependent-instruction latency, not by bytes.
template <int HG>
__global__ void __launch_bounds__(Cfg<HG>::THREADS, 640 / Cfg<HG>::THREADS) attn_bwd_fused_kernel(AttnArgs a) {
  constexpr int THREADS = Cfg<HG>::THREADS, LDS = Cfg<HG>::LDS, COLS = Cfg<HG>::COLS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);
  bf16* Gs = Qs + RC * LDS;      // dO
  bf16* Ks = Gs + RC * LDS;
  bf16* Vs = Ks + RC * LDS;
  float* Bs = reinterpret_cast<float*>(Vs + RC * LDS);   // [64] additive key bias
  float* Kp = Bs + RC;                                   // [64] 1 for kept keys else 0
  float4* St = reinterpret_cast<float4*>(Kp + RC);       // [64 queries][HG]: m, 1/l, D, -
  pdl_trigger();
  pdl_wait();
  const int d = blockIdx.z, b = blockIdx.y, hg = blockIdx.x;
  int qs_, nq;
  seg_patient(a.q, d, b, qs_, nq);     // this patient's (packed) query rows
  if (b == 0) zero_pad_slice<HG>(a.q, d, reinterpret_cast<bf16*>(a.dq) + hg * COLS, D);
  const int nk = a.kv.T[d];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hl = warp % HG, half = warp / HG;
  const int g = lane >> 2, t = lane & 3;
  const size_t qrow0 = (size_t)a.q.row0[d] + (size_t)qs_;
  const size_t krow0 = (size_t)a.kv.row0[d] + (size_t)b * nk;
  const bf16* kvsrc = reinterpret_cast<const bf16*>(a.kvbuf) + krow0 * a.ldkv + a.col0 + hg * COLS;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * nk : nullptr;
  const int nq16 = (nq + 15) & ~15, nk16 = (nk + 15) & ~15;
  stage<HG>(Qs, reinterpret_cast<const bf16*>(a.qb) + qrow0 * D + hg * COLS, D, nq, nq16);
  stage<HG>(Gs, reinterpret_cast<const bf16*>(a.d_o) + qrow0 * D + hg * COLS, D, nq, nq16);
  stage<HG>(Ks, kvsrc, a.ldkv, nk, nk16);
  stage<HG>(Vs, kvsrc + D, a.ldkv, nk, nk16);
  if (threadIdx.x < RC) {
    const float mk = threadIdx.x < nk ? (km ? (km[threadIdx.x] < 0.5f ? 0.f : 1.f) : 1.f) : -1.f;
    Bs[threadIdx.x] = key_bias(mk);
    Kp[threadIdx.x] = mk > 0.f ? 1.f : 0.f;
  }
  for (int idx = threadIdx.x; idx < RC * HG; idx += THREADS) {
    const int r = idx / HG, hh = idx % HG;
    float2 ml = make_float2(0.f, 0.f);     // 1/l = 0 for tile-padding queries -> P = 0
    if (r < nq) ml = *reinterpret_cast<const float2*>(a.ml + ((qrow0 + r) * H + hg * HG + hh) * 2);
    St[idx] = make_float4(ml.x, ml.y, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();
  const int NTK = (nk + 7) >> 3, NTQ = (nq + 7) >> 3;
  bf16* dq_out = reinterpret_cast<bf16*>(a.dq) + qrow0 * D + hg * COLS;

  // ---- pass A: rows = queries ------------------------------------------------------------
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nq) break;
    uint32_t qa[2][4], ga[2][4];
    frag_a<LDS>(Qs, mt * 16, hl * 32, lane, qa[0]);
    frag_a<LDS>(Qs, mt * 16, hl * 32 + 16, lane, qa[1]);
    frag_a<LDS>(Gs, mt * 16, hl * 32, lane, ga[0]);
    frag_a<LDS>(Gs, mt * 16, hl * 32 + 16, lane, ga[1]);
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const float4 st0 = St[r0 * HG + hl], st1 = St[r1 * HG + hl];
    uint32_t pp[8][2];          // P of this m-tile, packed bf16x2: [nt][0] = row g, [nt][1] = row g+8
    float dp[8][4];
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      pp[nt][0] = pp[nt][1] = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) dp[nt][j] = 0.f;
      if (nt < NTK) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t kb[4], vb[4];
        frag_b_nk<LDS>(Ks, nt * 8, hl * 32, lane, kb);
        frag_b_nk<LDS>(Vs, nt * 8, hl * 32, lane, vb);
        mma16816(s, qa[0], kb[0], kb[1]);
        mma16816(s, qa[1], kb[2], kb[3]);
        mma16816(dp[nt], ga[0], vb[0], vb[1]);
        mma16816(dp[nt], ga[1], vb[2], vb[3]);
        const float2 kbias = *reinterpret_cast<const float2*>(Bs + nt * 8 + 2 * t);
        const float p0 = ex2((rbf(s[0]) + kbias.x - st0.x) * L2E) * st0.y;
        const float p1 = ex2((rbf(s[1]) + kbias.y - st0.x) * L2E) * st0.y;
        const float p2 = ex2((rbf(s[2]) + kbias.x - st1.x) * L2E) * st1.y;
        const float p3 = ex2((rbf(s[3]) + kbias.y - st1.x) * L2E) * st1.y;
        D0 = fmaf(p0, dp[nt][0], fmaf(p1, dp[nt][1], D0));
        D1 = fmaf(p2, dp[nt][2], fmaf(p3, dp[nt][3], D1));
        pp[nt][0] = pack_bf16(p0, p1);
        pp[nt][1] = pack_bf16(p2, p3);
      }
    }
    D0 = quad_sum(D0); D1 = quad_sum(D1);
    if (t == 0) { St[r0 * HG + hl].z = D0; St[r1 * HG + hl].z = D1; }
    float dq[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (2 * kk >= NTK) continue;
      uint32_t dsa[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int nt = 2 * kk + e;
        const float2 keep = *reinterpret_cast<const float2*>(Kp + nt * 8 + 2 * t);   // tiles >= NTK: P = 0 already
        const float p0 = __uint_as_float(pp[nt][0] << 16), p1 = __uint_as_float(pp[nt][0] & 0xffff0000u);
        const float p2 = __uint_as_float(pp[nt][1] << 16), p3 = __uint_as_float(pp[nt][1] & 0xffff0000u);
        dsa[e * 2] = pack_bf16(p0 * (dp[nt][0] - D0) * keep.x, p1 * (dp[nt][1] - D0) * keep.y);
        dsa[e * 2 + 1] = pack_bf16(p2 * (dp[nt][2] - D1) * keep.x, p3 * (dp[nt][3] - D1) * keep.y);
      }
#pragma unroll
      for (int nc = 0; nc < 2; ++nc) {
        uint32_t kb[4];
        frag_b_kn<LDS>(Ks, kk * 16, hl * 32 + nc * 16, lane, kb);
        mma16816(dq[2 * nc], dsa, kb[0], kb[1]);
        mma16816(dq[2 * nc + 1], dsa, kb[2], kb[3]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) put_c_global(dq_out, D, mt * 16, hl * 32 + n * 8, lane, dq[n], nq);
  }
  __syncthreads();   // D of every (query, head) visible

  // ---- pass B: rows = keys (transposed tiles) ---------------------------------------------
  bf16* dkv_out = reinterpret_cast<bf16*>(a.dkv) + krow0 * a.ldkv + a.col0 + hg * COLS;
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int mt = half * 2 + i;
    if (mt * 16 >= nk) break;
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int j = 0; j < 4; ++j) { dk[n][j] = 0.f; dv[n][j] = 0.f; }
    uint32_t ka[2][4], va[2][4];
    frag_a<LDS>(Ks, mt * 16, hl * 32, lane, ka[0]);
    frag_a<LDS>(Ks, mt * 16, hl * 32 + 16, lane, ka[1]);
    frag_a<LDS>(Vs, mt * 16, hl * 32, lane, va[0]);
    frag_a<LDS>(Vs, mt * 16, hl * 32 + 16, lane, va[1]);
    const float kb0 = Bs[mt * 16 + g], kb1 = Bs[mt * 16 + g + 8];      // this thread's two key rows
    const float kp0 = Kp[mt * 16 + g], kp1 = Kp[mt * 16 + g + 8];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (2 * kk >= NTQ) continue;
      uint32_t pa[4], dsa[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int nt = 2 * kk + e;
        float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t qb[4], gb[4];
        frag_b_nk<LDS>(Qs, nt * 8, hl * 32, lane, qb);
        frag_b_nk<LDS>(Gs, nt * 8, hl * 32, lane, gb);
        mma16816(s, ka[0], qb[0], qb[1]);
        mma16816(s, ka[1], qb[2], qb[3]);
        mma16816(dp, va[0], gb[0], gb[1]);
        mma16816(dp, va[1], gb[2], gb[3]);
        const float4 sa = St[(nt * 8 + 2 * t) * HG + hl], sb = St[(nt * 8 + 2 * t + 1) * HG + hl];   // query columns
        const float p0 = ex2((rbf(s[0]) + kb0 - sa.x) * L2E) * sa.y, p1 = ex2((rbf(s[1]) + kb0 - sb.x) * L2E) * sb.y;
        const float p2 = ex2((rbf(s[2]) + kb1 - sa.x) * L2E) * sa.y, p3 = ex2((rbf(s[3]) + kb1 - sb.x) * L2E) * sb.y;
        pa[e * 2] = pack_bf16(p0, p1);
        pa[e * 2 + 1] = pack_bf16(p2, p3);
        dsa[e * 2] = pack_bf16(p0 * (dp[0] - sa.z) * kp0, p1 * (dp[1] - sb.z) * kp0);
        dsa[e * 2 + 1] = pack_bf16(p2 * (dp[2] - sa.z) * kp1, p3 * (dp[3] - sb.z) * kp1);
      }
#pragma unroll
      for (int nc = 0; nc < 2; ++nc) {
        uint32_t gb[4], qb[4];
        frag_b_kn<LDS>(Gs, kk * 16, hl * 32 + nc * 16, lane, gb);
        frag_b_kn<LDS>(Qs, kk * 16, hl * 32 + nc * 16, lane, qb);
        mma16816(dv[2 * nc], pa, gb[0], gb[1]);
        mma16816(dv[2 * nc + 1], pa, gb[2], gb[3]);
        mma16816(dk[2 * nc], dsa, qb[0], qb[1]);
        mma16816(dk[2 * nc + 1], dsa, qb[2], qb[3]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      put_c_global(dkv_out, a.ldkv, mt * 16, hl * 32 + n * 8, lane, dk[n], nk);
      put_c_global(dkv_out + D, a.ldkv, mt * 16, hl * 32 + n * 8, lane, dv[n], nk);
    }
  }
}

}  // namespace amma
}  // namespace mmr
