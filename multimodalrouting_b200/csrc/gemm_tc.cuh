// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//  gemm_tc_kernel   C[r, n] = sum_k A[a(r), k] * B[b(seg)+n, k]     both operands K-major
//  wgrad_tc_kernel  out[seg][m, n] += sum_r dY[r, m] * X[x(r), n]   both operands MN-major
//
// Warp 0 lane 0 drives TMA (cp.async.bulk.tensor, 128B swizzle) through a ring of mbarrier-guarded
// smem stages, warp 1 lane 0 issues tcgen05.mma with smem descriptors and commits completion to
// mbarriers, warps 2-5 drain the 128 x BN fp32 accumulator from TMEM (tcgen05.ld 32x32b).  The
// forward/data-gradient GEMM is persistent (one CTA per SM, two TMEM accumulators, store-only
// epilogue); the weight-gradient GEMM runs one CTA per (tile, split-K chunk), two CTAs per SM.
#pragma once
#include <cuda.h>

#include "epilogue.cuh"

namespace mmr {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;       // 64 bf16 = one 128-byte swizzle row
constexpr int NTHREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must fault the launch (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (count 1) on the mbarrier once all previously issued tcgen05.mma of this thread retire
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- CTA-pair (cta_group::2) primitives: two SMs of a TPC cooperate on one 256-row tile ------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// Remote arrive on the peer CTA's barrier.  Default semantics (.release at .cta scope), as CUTLASS's ClusterBarrier::arrive
// does: the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of every arrive, which was 15 % of
// the chained kernel's stall samples (profiles/r2_chain.md).  What the arrive has to order is covered elsewhere: TMEM reads
// by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, shared-memory writes for the tensor core by fence.proxy.async.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the mbarrier at `bar`
// (a shared::cluster address -- the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at the same offset in BOTH CTAs of the pair once the issued MMAs retire
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// 16-byte vector reduction into global memory (split-K accumulation of weight gradients)
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): 128B swizzle, version 1.
//  K-major : 8-row groups of 128B rows, SBO = 1024 B, LBO unused.
//  MN-major: 64-element (128B) MN atoms x 8 k-rows; SBO = 1024 B between k groups,
//            LBO = 8192 B between MN atoms (each atom block is [64 k][64 mn] = 8 KB).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> fp32, M=128.
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BN> __host__ __device__ constexpr int stage_bytes() { return (BM + BN) * BK * 2; }
template <int BN> __host__ __device__ constexpr int smem_bytes(int stages) { return stages * stage_bytes<BN>() + 1024 + 512; }

struct Ctrl {   // lives after the stages
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t acc_full;      // wgrad kernel (single accumulator)
  uint64_t tfull[2];      // persistent GEMM: accumulator buffer ready for the epilogue
  uint64_t tempty[2];     // persistent GEMM: accumulator buffer drained
  uint32_t tmem_base;
  uint32_t pad_;
  // dynamic tile scheduling (cluster launch control): 16-byte try_cancel responses, one mbarrier pair per slot
  uint64_t clc_full[4];
  uint64_t clc_empty[4];
  uint4 clc_resp[4];
};
static_assert(sizeof(Ctrl) <= 512, "Ctrl must fit its shared-memory reserve");

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Epilogue of the persistent tcgen05 GEMM.  Everything that needs a global LOAD is row- or
// column-uniform (bias per column, keep-mask per row, 32 ReLU-mask bits per row and 32-column chunk)
// and is fetched once per tile before the accumulator is ready; after the shared-memory transpose the
// epilogue only stores, so no memory latency sits on the per-chunk critical path.
enum TcOp {
  TEPI_BIAS = 0,            // out<bf16> = acc + bias
  TEPI_BIAS_RELU_BITS = 1,  // out<bf16> = relu(acc + bias); bits_out = (acc + bias > 0)
  TEPI_BITS_IN = 2,         // out<bf16> = acc * bits_in
  TEPI_MASK = 3,            // out<bf16> = acc * rowmask   (rowmask may be null)
  TEPI_F32 = 4,             // out<f32>  = acc
  TEPI_BIAS_F32 = 5,        // out<f32>  = acc + bias      (bias may be null; unit tests)
};
struct TcEpi {
  const float* bias;          // stacked like the B rows (b_row0[seg] + n), or null
  const float* rowmask;       // [rows] in C row space, or null
  const uint32_t* bits_in;    // [rows, ld_bits] words; word n/32 holds columns n..n+31
  uint32_t* bits_out;
  int ld_bits;
  void* out;
  int ldo;
  int out_rows;               // rows physically present in `out` (TMA store bound); 0 = the padded row space
  int dbg;                    // tuning experiments only (MMR_TC_DBG, tools/bench_gemm.py): 1 = epilogue drains TMEM but neither
                              // converts nor stores, 2 = no TMA / bit stores, 4 = epilogue only hands the buffer back
};

constexpr int GEMM_THREADS = 320;     // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int STG_FLOATS = 32 * 36;   // per-warp transpose tile of the weight-gradient epilogue

// TMA-store staging per epilogue warp: one 128B-swizzled [32 rows x 128 bytes] box for bf16 outputs (64 columns per round),
// two for fp32 outputs (2 x 32 columns per round).  Round 1 staged a warp's whole 128-column slice (8 KB); halving it buys a
// fifth operand stage for the CTA-pair kernels -- these GEMMs are bound by bytes in flight per SM (profiles/r2_gemm_experiments.md).
template <bool F32> __host__ __device__ constexpr int cstg_bytes() { return F32 ? 8192 : 4096; }

// TMA store of one [32 rows x 128 bytes] box from shared memory (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// One 32-column chunk of one accumulator row: apply the epilogue op and write it into the warp's
// staging box (row = lane, 128-byte rows, 16-byte units XOR-swizzled by row & 7 like the TMA map).
//   bf16: `unit0` = first 16-byte unit of the chunk inside its box (0 or 4), 4 units
//   fp32: the chunk is a whole box row (8 units)
// relu(a), relu(b) rounded to bf16 and packed (lo = a): one F2FP instruction for two elements
__device__ __forceinline__ uint32_t pack2_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// wb |= (v > 0) << j   as FSETP + predicated LOP3
__device__ __forceinline__ void or_bit_if_pos(uint32_t& wb, float v, uint32_t bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(wb) : "f"(v), "r"(bit));
}

template <int OP>
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], const float* bias_c, float rmask, uint32_t wbits_in,
                                          uint32_t& wbits_out, bool zero_row, uint32_t box_row_addr, int unit0, int lane) {
  // `zero_row`: this thread's row is segment padding (only possible in the last tile of a segment; the caller
  // passes a tile-uniform false otherwise so the selects below fold away at run time)
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (OP == TEPI_BIAS || OP == TEPI_BIAS_RELU_BITS || OP == TEPI_BIAS_F32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(bias_c + j);   // shared-memory broadcast
      v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
    }
  }
  if (OP == TEPI_BIAS_RELU_BITS) {
    uint32_t wb = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) or_bit_if_pos(wb, v[j], 1u << j);
    wbits_out = wb;          // the ReLU itself is folded into the bf16 pack below
  }
  if (OP == TEPI_BITS_IN) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = ((wbits_in >> j) & 1u) ? v[j] : 0.f;
  }
  if (OP == TEPI_MASK) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= rmask;
  }
  if (zero_row) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  const uint32_t sw = (uint32_t)(lane & 7);
  if (OP == TEPI_F32 || OP == TEPI_BIAS_F32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t addr = box_row_addr + (((uint32_t)j ^ sw) << 4);
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t addr = box_row_addr + (((uint32_t)(unit0 + j) ^ sw) << 4);
      uint32_t w0, w1, w2, w3;
      if (OP == TEPI_BIAS_RELU_BITS) {
        w0 = pack2_bf16_relu(v[8 * j], v[8 * j + 1]); w1 = pack2_bf16_relu(v[8 * j + 2], v[8 * j + 3]);
        w2 = pack2_bf16_relu(v[8 * j + 4], v[8 * j + 5]); w3 = pack2_bf16_relu(v[8 * j + 6], v[8 * j + 7]);
      } else {
        w0 = pack2_bf16(v[8 * j], v[8 * j + 1]); w1 = pack2_bf16(v[8 * j + 2], v[8 * j + 3]);
        w2 = pack2_bf16(v[8 * j + 4], v[8 * j + 5]); w3 = pack2_bf16(v[8 * j + 6], v[8 * j + 7]);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------
// Persistent, warp-specialised GEMM: grid = #SMs, each CTA walks 128 x BN output tiles.
//   warp 0    : TMA producer, runs ahead across tile boundaries through the smem ring
//   warp 1    : tcgen05.mma issuer, alternates between two BN-column TMEM accumulators
//   warps 2-9 : epilogue; warp w drains TMEM lane quarter (w & 3), column half ((w - 2) >> 2) of
//               accumulator i while the MMA warp fills accumulator i+1
// CL = 2: CTA pairs (cluster of two SMs, tcgen05 cta_group::2).  The pair owns a 256 x BN tile; each CTA
// stages its own 128 A rows and only HALF of the B tile (BN/2 weight rows), the leader CTA's single thread
// issues M=256 MMAs that read both CTAs' shared memory and write both CTAs' TMEM.  Operand bytes entering
// each SM per k-block drop from 48 KB to 32 KB (the measured bound, profiles/r1_gemm_epilogue.md) and a
// fourth stage fits.  Requires every segment to hold an even number of 128-row tiles.
template <int BN, int CL> __host__ __device__ constexpr int gemm_stage_bytes() { return (BM + BN / CL) * BK * 2; }
template <int BN, int CL, bool F32> __host__ __device__ constexpr int gemm_smem_bytes(int stages) {
  return stages * gemm_stage_bytes<BN, CL>() + 8 * cstg_bytes<F32>() + 8 * (BN / 2) * 4 + 1024 + 512;
}
// DYN = 1 (CTA pairs only): dynamic tile scheduling through cluster launch control.  The grid holds one cluster per
// (static) work item; a resident cluster starts with its own item and then CANCELS pending clusters of the grid
// (clusterlaunchcontrol.try_cancel) and takes over theirs, so the items go to whichever pairs actually hold SMs -- a pair
// that starts late because the weight-gradient stream or NCCL occupies its SMs simply takes fewer.  The 16-byte response is
// multicast by the hardware to the same shared-memory offset of both CTAs and completes an mbarrier in each (clc_full);
// every role reads the item from there and acknowledges on the leader's clc_empty before the slot is reused.  The leader's
// MMA thread is the scheduler: at the start of item i it looks at the response for item i + 1 and, if that was a
// cancellation, requests item i + 2.  Items beyond the valid (device-side) count are skipped by every role alike.
template <int BN, int OP, int CL, int DYN = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, GemmProblem g, TcEpi e, int stages) {
  static_assert(DYN == 0 || CL == 2, "dynamic scheduling is implemented for CTA pairs");
  constexpr int STAGE = gemm_stage_bytes<BN, CL>();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t cstg_base = sbase + stages * STAGE;    // 1024-byte aligned (stages are multiples of 1 KB)
  constexpr int CSTG_BYTES = cstg_bytes<OP == TEPI_F32 || OP == TEPI_BIAS_F32>();
  float* bias_all = reinterpret_cast<float*>(sgen + stages * STAGE + 8 * CSTG_BYTES);
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(bias_all + 8 * (BN / 2));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL == 2 ? cluster_ctarank() : 0u;
  const int kblocks = g.K / BK;
  const int nN = g.N / BN;
  // work items: 128-row tiles (CL = 1) or 256-row tile pairs (CL = 2), n-block fastest
  // Work items = the 128-row tiles (CL = 1) / 256-row tile pairs (CL = 2) that hold valid rows, n-block fastest, enumerated
  // segment by segment.  With packed query rows the valid-row counts live in device memory; every role (producer, MMA issuer,
  // epilogue warps) derives the same list, so the stage / accumulator counters stay in step and the static round-robin over the
  // CTAs is balanced to within one item.
  int vt0[7];
  vt0[0] = 0;
#pragma unroll
  for (int sgi = 0; sgi < 6; ++sgi) {
    const int rows_s = sgi < g.segs.n ? (g.segs.nv ? seg_rows(g.segs, sgi) : g.segs.row0[sgi + 1] - g.segs.row0[sgi]) : 0;
    vt0[sgi + 1] = vt0[sgi] + (rows_s + CL * BM - 1) / (CL * BM);
  }
  const int nwork = vt0[6] * nN;
  const int w0 = blockIdx.x / CL, wstep = gridDim.x / CL;
  // first row (of the tile / tile pair) of work item t
  auto item_row = [&](int t) -> int {
    const int mt = t / nN;
    int sg = 0;
#pragma unroll
    for (int i = 1; i < 6; ++i)
      if (mt >= vt0[i]) sg = i;
    return g.segs.row0[sg] + (mt - vt0[sg]) * CL * BM;
  };

  // item i >= 1 of this cluster travels through slot (i - 1) & 3; its k-th use completes phase k of the slot's barriers
  auto clc_issue = [&](uint32_t j) {      // leader's scheduler thread
    const uint32_t sl = (j - 1) & 3u, k = (j - 1) >> 2;
    mbar_wait(smem_u32(&ctrl->clc_empty[sl]), (k & 1u) ^ 1u);      // every role of both CTAs has read the previous use
    const uint32_t fb = smem_u32(&ctrl->clc_full[sl]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(fb) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], 16;" ::"r"(mapa_u32(fb, 1)) : "memory");
    asm volatile(
        "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
        ::"r"(smem_u32(&ctrl->clc_resp[sl])), "r"(fb) : "memory");
  };
  // item index of this cluster's i-th item, -1 when the grid has no pending cluster left; `ack`: this caller is one of the
  // 18 acknowledging readers (producer lanes, epilogue warps), the scheduler thread reads without acknowledging
  auto clc_item = [&](uint32_t i, bool ack) -> int {
    if (i == 0) return (int)(blockIdx.x / CL);
    const uint32_t sl = (i - 1) & 3u, k = (i - 1) >> 2;
    mbar_wait(smem_u32(&ctrl->clc_full[sl]), k & 1u);
    uint32_t valid = 0, x = 0, y = 0, z = 0;
    asm volatile(
        "{\n\t.reg .pred p1;\n\t.reg .b128 clc_result;\n\t"
        "ld.shared.b128 clc_result, [%4];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, clc_result;\n\t"
        "selp.u32 %3, 1, 0, p1;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, clc_result;\n\t}"
        : "+r"(x), "+r"(y), "+r"(z), "+r"(valid) : "r"(smem_u32(&ctrl->clc_resp[sl])) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the slot is rewritten through the async proxy
    if (ack) {
      if (warp >= 2) __syncwarp();        // epilogue warps: every lane has read the slot (the producers call with one lane)
      if (lane == 0) {
        if (rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&ctrl->clc_empty[sl]), 0));
        else mbar_arrive(smem_u32(&ctrl->clc_empty[sl]));
      }
    }
    return valid ? (int)(x / CL) : -1;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&ctrl->full[s]), 1);
      mbar_init(smem_u32(&ctrl->empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&ctrl->tfull[b]), 1);
      mbar_init(smem_u32(&ctrl->tempty[b]), 8 * CL);
    }
    if (DYN) {
      for (int s = 0; s < 4; ++s) {
        mbar_init(smem_u32(&ctrl->clc_full[s]), 1);
        mbar_init(smem_u32(&ctrl->clc_empty[s]), 18);     // 2 producer lanes + 16 epilogue warps
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CL == 2) tmem_alloc_pair(smem_u32(&ctrl->tmem_base), 2 * BN);
    else tmem_alloc(smem_u32(&ctrl->tmem_base), 2 * BN);
  }
  tc_fence_before();
  if (CL == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_wait();      // barriers / TMEM are set up; operands of the previous kernel are complete from here on

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
      uint32_t it = 0;
      for (uint32_t wi = 0;; ++wi) {
        int t;
        if (DYN) {
          t = clc_item(wi, true);
          if (t < 0) break;
          if (t >= nwork) continue;
        } else {
          t = w0 + (int)wi * wstep;
          if (t >= nwork) break;
        }
        const int m0 = item_row(t) + (int)rank * BM, n0 = (t % nN) * BN;
        const int seg = seg_of_row(g.segs, m0);
        const int a_row = g.a_row0[seg] + (m0 - g.segs.row0[seg]);
        const int b_row = g.b_row0[seg] + n0 + (int)rank * (BN / CL);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(smem_u32(&ctrl->empty[s]), ph ^ 1);
          const uint32_t sa = sbase + s * STAGE;
          const uint32_t sb = sa + BM * BK * 2;
          if (CL == 2) {
            // both CTAs credit the LEADER's barrier (peer bit of the shared::cluster address cleared)
            const uint32_t full = mapa_u32(smem_u32(&ctrl->full[s]), 0);
            if (rank == 0) mbar_expect_tx(smem_u32(&ctrl->full[s]), 2 * STAGE);
            tma_load_2d_pair(sa, &tmA, kb * BK, a_row, full);
            tma_load_2d_pair(sb, &tmB, kb * BK, b_row, full);
          } else {
            const uint32_t full = smem_u32(&ctrl->full[s]);
            mbar_expect_tx(full, STAGE);
            tma_load_2d(sa, &tmA, kb * BK, a_row, full);
            tma_load_2d(sb, &tmB, kb * BK, b_row, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc(BN, 0, 0, CL * BM);
      uint32_t it = 0, i = 0;
      if (DYN) clc_issue(1);
      for (uint32_t wi = 0;; ++wi) {
        int t;
        if (DYN) {
          t = clc_item(wi, false);
          if (t < 0) break;
          if (clc_item(wi + 1, false) >= 0) clc_issue(wi + 2);     // never request after a failed cancellation
          if (t >= nwork) continue;
        } else {
          t = w0 + (int)wi * wstep;
          if (t >= nwork) break;
        }
        const uint32_t buf = i & 1;
        mbar_wait(smem_u32(&ctrl->tempty[buf]), ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(smem_u32(&ctrl->full[s]), ph);
          tc_fence_after();
          const uint32_t sa = sbase + s * STAGE;
          const uint32_t sb = sa + BM * BK * 2;
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {   // +32 bytes (>>4 = 2) per K=16 step inside the swizzle row
            if (CL == 2) umma_bf16_pair(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            else umma_bf16(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          if (CL == 2) umma_commit_pair(smem_u32(&ctrl->empty[s]));
          else umma_commit(smem_u32(&ctrl->empty[s]));
        }
        if (CL == 2) umma_commit_pair(smem_u32(&ctrl->tfull[buf]));
        else umma_commit(smem_u32(&ctrl->tfull[buf]));
        ++i;
      }
    }
  } else {
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;       // column half
    constexpr int HC = BN / 2;            // columns per epilogue warp
    float* bias_s = bias_all + (warp - 2) * HC;
    const bool f32_out = (OP == TEPI_F32 || OP == TEPI_BIAS_F32);
    uint32_t i = 0;
    for (uint32_t wi = 0;; ++wi) {
      int t;
      if (DYN) {
        t = clc_item(wi, true);
        if (t < 0) break;
        if (t >= nwork) continue;
      } else {
        t = w0 + (int)wi * wstep;
        if (t >= nwork) break;
      }
      const uint32_t i_cur = i++;
      const int m0 = item_row(t) + (int)rank * BM, n0 = (t % nN) * BN + ch * HC;
      const int seg = seg_of_row(g.segs, m0);
      const int rows_valid = seg_rows(g.segs, seg) - (m0 - g.segs.row0[seg]);
      const int lr = q * 32 + lane;                 // accumulator row owned by this thread
      const bool row_ok = lr < rows_valid;
      const bool zrow = (rows_valid < BM) && !row_ok;   // tile-uniform fast path when the tile has no padding rows
      // ---- per-tile operands, fetched while the accumulator is still being computed ----
      float rmask = 1.f;
      uint32_t bits[HC / 32];
      if (OP == TEPI_MASK) rmask = (e.rowmask != nullptr && row_ok) ? e.rowmask[m0 + lr] : 1.f;
      if (OP == TEPI_BITS_IN) {
        const uint4* bp = reinterpret_cast<const uint4*>(e.bits_in + (size_t)(m0 + lr) * e.ld_bits + n0 / 32);
#pragma unroll
        for (int w4 = 0; w4 < HC / 128; ++w4) {
          uint4 b4 = row_ok ? bp[w4] : make_uint4(0, 0, 0, 0);
          bits[4 * w4] = b4.x; bits[4 * w4 + 1] = b4.y; bits[4 * w4 + 2] = b4.z; bits[4 * w4 + 3] = b4.w;
        }
      }
      if (OP == TEPI_BIAS || OP == TEPI_BIAS_RELU_BITS || OP == TEPI_BIAS_F32) {
        const float* bsrc = e.bias ? e.bias + g.b_row0[seg] + n0 : nullptr;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < HC / 32; ++j) bias_s[lane + 32 * j] = bsrc ? bsrc[lane + 32 * j] : 0.f;
        __syncwarp();
      }
      const uint32_t buf = i_cur & 1;
      mbar_wait(smem_u32(&ctrl->tfull[buf]), (i_cur >> 1) & 1);
      tc_fence_after();
      if (e.dbg & 4) {           // experiment: no drain at all
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&ctrl->tempty[buf]), 0));
          else mbar_arrive(smem_u32(&ctrl->tempty[buf]));
        }
        continue;
      }
      const uint32_t tmem_acc = tmem_base + buf * BN + ch * HC + ((uint32_t)(q * 32) << 16);
      const uint32_t stg = cstg_base + (warp - 2) * CSTG_BYTES;     // one (bf16) / two (fp32) 4 KB boxes
      const uint32_t stg_row = stg + lane * 128;
      uint32_t wout[HC / 32];
#pragma unroll
      for (int cc = 0; cc < HC / 64; ++cc) {
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(tmem_acc + cc * 64, r0);
        tmem_ld32_nowait(tmem_acc + cc * 64 + 32, r1);
        tmem_ld_wait();
        if (cc == HC / 64 - 1) {   // accumulator fully read: hand the TMEM buffer back to the (leader's) MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CL == 2 && rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&ctrl->tempty[buf]), 0));
            else mbar_arrive(smem_u32(&ctrl->tempty[buf]));
          }
        }
        if (e.dbg & 1) continue;   // experiment: TMEM drained, nothing converted or stored
        // the staging box(es) may still be read by the previous TMA store of this warp
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        if (f32_out) {   // 64 fp32 columns = two [32 x 32] boxes per round
          epi_chunk<OP>(r0, bias_s + cc * 64, rmask, 0u, wout[2 * cc], zrow, stg_row, 0, lane);
          epi_chunk<OP>(r1, bias_s + cc * 64 + 32, rmask, 0u, wout[2 * cc + 1], zrow, stg_row + 4096, 0, lane);
        } else {         // 64 bf16 columns = one [32 x 64] box per round
          epi_chunk<OP>(r0, bias_s + cc * 64, rmask, OP == TEPI_BITS_IN ? bits[2 * cc] : 0u, wout[2 * cc], zrow,
                        stg_row, 0, lane);
          epi_chunk<OP>(r1, bias_s + cc * 64 + 32, rmask, OP == TEPI_BITS_IN ? bits[2 * cc + 1] : 0u, wout[2 * cc + 1],
                        zrow, stg_row, 4, lane);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && !(e.dbg & 2)) {
          tma_store_2d(&tmC, stg, n0 + cc * 64, m0 + q * 32);
          if (f32_out) tma_store_2d(&tmC, stg + 4096, n0 + cc * 64 + 32, m0 + q * 32);
          tma_store_commit();
        }
      }
      if (OP == TEPI_BIAS_RELU_BITS && row_ok && !(e.dbg & 3)) {
        uint4* bo = reinterpret_cast<uint4*>(e.bits_out + (size_t)(m0 + lr) * e.ld_bits + n0 / 32);
#pragma unroll
        for (int w4 = 0; w4 < HC / 128; ++w4) bo[w4] = make_uint4(wout[4 * w4], wout[4 * w4 + 1], wout[4 * w4 + 2], wout[4 * w4 + 3]);
      }
    }
  }
  if (warp >= 2 && lane == 0) tma_store_wait_all();   // bulk stores must complete before the CTA exits
  tc_fence_before();
  if (CL == 2) cluster_sync_all(); else __syncthreads();   // the peer may still read this CTA's smem / arrive on its barriers
  if (warp == 1) {
    if (CL == 2) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient: reduction over token rows; A = dY^T and B = X^T are MN-major views of the
// row-major activations, fetched as [64 rows][64 cols] TMA boxes.  grid.z = nseg * splits.
// MT = 128-row output tiles per CTA: MT = 2 (a 256 x 256 tile, two TMEM accumulators fed by the same X stage)
// cuts the L2 -> SM operand traffic by a third -- the measured bound of this kernel (profiles/r1_wgrad.md).
// split-K chunks per segment (proportional to the segment's rows so that all CTAs reduce similar row counts):
// blockIdx.z in [first[seg], first[seg+1]) works on segment seg
struct WgradSplits { int first[7]; };

template <int MT, int BN> __host__ __device__ constexpr int wgrad_stage_bytes() { return (MT * BM + BN) * BK * 2; }
template <int MT, int BN> __host__ __device__ constexpr int wgrad_smem_bytes(int stages) {
  return stages * wgrad_stage_bytes<MT, BN>() + 1024 + 512;
}

template <int MT, int BN>
__global__ void __launch_bounds__(NTHREADS, MT == 1 ? 2 : 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX,
                WgradProblem w, WgradSplits sp_tab, int stages) {
  constexpr int STAGE = wgrad_stage_bytes<MT, BN>();
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  Ctrl* ctrl = reinterpret_cast<Ctrl*>(sgen + stages * STAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int seg = 0, sp, splits;
  if (sp_tab.first[0] < 0) {
    // Balanced split-K from the DEVICE-side row counts (packed query rows: the valid rows of a segment are data dependent,
    // 4.3 k to 25 k at B = 512, while a static table gives every segment the same number of chunks): the gridDim.z chunks
    // are dealt to the segments in proportion to their k-blocks -- one each, the rest by floor, leftovers to the segment
    // with the longest chunks -- so every CTA reduces about total / gridDim.z k-blocks.  Every thread computes the same table.
    const int Z = gridDim.z, n = w.segs.n;
    int kb[6], cnt[6], total = 0, nonempty = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      kb[i] = 0; cnt[i] = 0;
      if (i < n) {
        kb[i] = w.segs.nv ? (seg_rows(w.segs, i) + BK - 1) / BK : (w.segs.row0[i + 1] - w.segs.row0[i]) / BK;
        total += kb[i];
        if (kb[i] > 0) { cnt[i] = 1; ++nonempty; }
      }
    }
    int rest = Z - nonempty, used = 0;
    if (total > 0 && rest > 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const int add = (int)(((long long)rest * kb[i]) / total);
        cnt[i] += add; used += add;
      }
      for (int left = rest - used; left > 0; --left) {      // at most n iterations
        int best = 0; long long bv = -1;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const long long v = cnt[i] > 0 ? ((long long)kb[i] << 20) / cnt[i] : -1;
          if (v > bv) { bv = v; best = i; }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i == best) ++cnt[i];
      }
    }
    int first = 0, z = blockIdx.z;
    sp = -1; splits = 1;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (sp < 0 && z < first + cnt[i]) { seg = i; sp = z - first; splits = cnt[i]; }
      first += cnt[i];
    }
    if (sp < 0) return;       // more chunks than work (tiny batches)
  } else {
#pragma unroll
    for (int i = 1; i < 6; ++i)
      if (i < w.segs.n && (int)blockIdx.z >= sp_tab.first[i]) seg = i;
    sp = blockIdx.z - sp_tab.first[seg]; splits = sp_tab.first[seg + 1] - sp_tab.first[seg];
  }
  const int m0 = blockIdx.y * (MT * BM), n0 = blockIdx.x * BN;
  // reduction rows: the whole padded segment, or (packed query rows) the valid rows rounded up to a k-block -- rows
  // [nv, pad256(nv)) are zero in both operands by the q-space contract
  const int rows_pad = w.segs.nv ? (seg_rows(w.segs, seg) + BK - 1) / BK * BK : w.segs.row0[seg + 1] - w.segs.row0[seg];
  const int kb_total = rows_pad / BK;
  const int kb_per = (kb_total + splits - 1) / splits;
  const int kb_begin = sp * kb_per;
  const int kb_end = min(kb_total, kb_begin + kb_per);
  const int kblocks = kb_end - kb_begin;
  if (kblocks <= 0) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&ctrl->full[s]), 1);
      mbar_init(smem_u32(&ctrl->empty[s]), w.colsum ? 5 : 1);   // MMA commit (+ the 4 column-sum warps)
    }
    mbar_init(smem_u32(&ctrl->acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(&ctrl->tmem_base), MT * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = ctrl->tmem_base;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmY)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
      for (int i = 0; i < kblocks; ++i) {
        const int s = i % stages;
        const uint32_t ph = (i / stages) & 1;
        mbar_wait(smem_u32(&ctrl->empty[s]), ph ^ 1);
        const uint32_t full = smem_u32(&ctrl->full[s]);
        mbar_expect_tx(full, STAGE);
        const uint32_t sa = sbase + s * STAGE;
        const uint32_t sb = sa + MT * BM * BK * 2;
        const int ry = w.segs.row0[seg] + (kb_begin + i) * BK;
        const int rx = w.x_row0[seg] + (kb_begin + i) * BK;
#pragma unroll
        for (int a = 0; a < MT * BM / 64; ++a) tma_load_2d(sa + a * 8192, &tmY, m0 + a * 64, ry, full);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b) tma_load_2d(sb + b * 8192, &tmX, n0 + b * 64, rx, full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, 1, 1);
      for (int i = 0; i < kblocks; ++i) {
        const int s = i % stages;
        const uint32_t ph = (i / stages) & 1;
        mbar_wait(smem_u32(&ctrl->full[s]), ph);
        tc_fence_after();
        const uint32_t sa = sbase + s * STAGE;
        const uint32_t sb = sa + MT * BM * BK * 2;
        const uint64_t bdesc = make_smem_desc(sb, 8192, 1024);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint64_t adesc = make_smem_desc(sa + mt * (BM * BK * 2), 8192, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // 16 k-rows of 128 B = 2048 B (>>4 = 128) per K=16 step
            umma_bf16(tmem_acc + mt * BN, adesc + 128 * k, bdesc + 128 * k, idesc, (i | k) != 0);
        }
        umma_commit(smem_u32(&ctrl->empty[s]));
      }
      umma_commit(smem_u32(&ctrl->acc_full));
    }
  } else {
    const int q = warp & 3;
    if (w.colsum) {
      // Bias gradient fused in: while the MMA warp consumes the stages, the (otherwise idle) epilogue warps
      // sum the dY tile over its 64 reduction rows.  Warp q takes rows [16q, 16q+16) of every stage, lane l the
      // 4 columns 4l..4l+3 of each 128-column group (atom l/16, 16-byte unit (l%16)/2 of the 128B-swizzled
      // [64 k][64 m] atom).
      const bool mine = blockIdx.x == 0 && w.dbias[seg] != nullptr;
      float cs[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) cs[mt][0] = cs[mt][1] = cs[mt][2] = cs[mt][3] = 0.f;
      const uint32_t unit = (lane & 15) >> 1, sub = (lane & 1) * 8, atom = lane >> 4;
      for (int i = 0; i < kblocks; ++i) {
        const int s = i % stages;
        mbar_wait(smem_u32(&ctrl->full[s]), (i / stages) & 1);
        if (mine) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint32_t sa = sbase + s * STAGE + (mt * 2 + atom) * 8192;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
              const uint32_t k = q * 16 + kk;
              uint32_t lo, hi;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(sa + k * 128 + ((unit ^ (k & 7)) << 4) + sub));
              cs[mt][0] += __uint_as_float(lo << 16); cs[mt][1] += __uint_as_float(lo & 0xffff0000u);
              cs[mt][2] += __uint_as_float(hi << 16); cs[mt][3] += __uint_as_float(hi & 0xffff0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ctrl->empty[s]));
      }
      if (mine) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int mcol = m0 + mt * BM + 4 * lane;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (mcol + j < w.M) atomicAdd(w.dbias[seg] + mcol + j, cs[mt][j]);
        }
      }
    }
    mbar_wait(smem_u32(&ctrl->acc_full), 0);
    tc_fence_after();
    float* out = w.out[seg];
    float* stg = reinterpret_cast<float*>(sgen) + q * (32 * 36);
    const int rg = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
    for (int c = 0; c < MT * BN / 32; ++c) {
      const int mt = c / (BN / 32), cn = c % (BN / 32);
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + c * 32, v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int n = n0 + cn * 32 + c4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + mt * BM + q * 32 + 4 * i + rg;
        const float4 a4 = *reinterpret_cast<const float4*>(stg + (4 * i + rg) * 36 + c4);
        if (out != nullptr && m < w.M && n < w.N) red_add_v4(out + (size_t)m * w.ldo + n, a4);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_acc, MT * BN);
}

// ---------------------------------------------------------------------------- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2-D row-major tensor [outer, inner] (inner contiguous, leading dimension ld elements), bf16 or fp32.
static bool make_tmap(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t ld,
                      uint32_t box_inner, uint32_t box_outer, bool f32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}
// programmatic dependent launch of the per-layer kernels (MMR_PDL=0 disables)
inline bool pdl_enabled() {
  static int on = env_int("MMR_PDL", 0);   // measured: -2.5 % at B=512 (early CTAs compete with the tail), so opt-in
  return on != 0;
}

template <int OP, int CL, int DYN = 0>
static cudaError_t launch_gemm_tc_cl(const GemmProblem& g, const TcEpi& e, int a_rows_total, int b_rows_total,
                                     int sm_count, cudaStream_t st) {
  constexpr int BN = 256;
  constexpr bool f32_stg = (OP == TEPI_F32 || OP == TEPI_BIAS_F32);
  // 32 KB (pair) / 48 KB operand stages next to 32 KB (bf16) / 64 KB (fp32) of TMA-store staging
  const int max_stages = CL == 2 ? (f32_stg ? 4 : 5) : (f32_stg ? 3 : 3);
  static int stages_cfg = env_int("MMR_TC_STAGES", 8);
  const int stages = stages_cfg < 1 ? 1 : (stages_cfg > max_stages ? max_stages : stages_cfg);
  constexpr bool f32_out = (OP == TEPI_F32 || OP == TEPI_BIAS_F32);
  CUtensorMap tmA, tmB, tmC;
  const uint64_t c_rows = e.out_rows > 0 ? (uint64_t)e.out_rows : (uint64_t)g.segs.row0[g.segs.n];   // TMA clips rows beyond
  if (!make_tmap(&tmC, e.out, (uint64_t)e.ldo, c_rows, (uint64_t)e.ldo, f32_out ? 32 : 64, 32, f32_out))
    return cudaErrorUnknown;
  if (!make_tmap(&tmA, g.A, (uint64_t)g.K, (uint64_t)a_rows_total, (uint64_t)g.lda, BK, BM)) return cudaErrorUnknown;
  if (!make_tmap(&tmB, g.B, (uint64_t)g.K, (uint64_t)b_rows_total, (uint64_t)g.ldb, BK, BN / CL)) return cudaErrorUnknown;
  auto kern = gemm_tc_kernel<BN, OP, CL, DYN>;
  const int smem = gemm_smem_bytes<BN, CL, f32_stg>(stages);
  static int dbg_cfg = env_int("MMR_TC_DBG", 0);
  TcEpi e2 = e;
  e2.dbg = dbg_cfg;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return err;
  const int total_rows = g.segs.row0[g.segs.n];
  const int nwork = ((total_rows + CL * BM - 1) / (CL * BM)) * (g.N / BN);
  int grid = nwork * CL < sm_count ? nwork * CL : sm_count;
  grid -= grid % CL;
  if (DYN) grid = nwork * CL;      // one cluster per static work item; resident clusters cancel the pending ones
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, g, e2, stages);
}

// a_rows_total / b_rows_total: number of rows physically present in A / B (TMA bounds).
template <int OP>
static cudaError_t launch_gemm_tc(const GemmProblem& g, const TcEpi& e, int a_rows_total, int b_rows_total,
                                  cudaStream_t st) {
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  // CTA pairs need every segment to start on a 256-row boundary and hold an even number of 128-row tiles
  // (measured, tools/bench_gemm.py, profiles/r2_gemm_experiments.md: with the cheap remote arrive pairs win on every shape
  // of the path -- K = 256: q/out 28.8 -> 27.1 us, fc1 82.2 -> 75.8 us)
  static int pair_cfg = env_int("MMR_TC_PAIR", 1);
  static int pair_min_k = env_int("MMR_TC_PAIR_MIN_K", 256);
  bool pair_ok = pair_cfg != 0 && g.K >= pair_min_k;
  for (int s = 0; s <= g.segs.n && pair_ok; ++s) pair_ok = g.segs.row0[s] % (2 * BM) == 0;
  // dynamic tile scheduling through cluster launch control: opt-in (read per launch: the tests switch it inside one process).
  // Measured (profiles/r2_experiments_late.md): parity green, the GEMM class alone +1.3 %, the step equal at N = 1, and the
  // overlapped all-reduce at N = 2 equally slow with it -- the static schedule was not what made the overlap lose.
  const int clc_cfg = env_int("MMR_TC_CLC", 0);
  if (pair_ok && clc_cfg) return launch_gemm_tc_cl<OP, 2, 1>(g, e, a_rows_total, b_rows_total, sm_count, st);
  if (pair_ok) return launch_gemm_tc_cl<OP, 2>(g, e, a_rows_total, b_rows_total, sm_count, st);
  return launch_gemm_tc_cl<OP, 1>(g, e, a_rows_total, b_rows_total, sm_count, st);
}

template <int MT>
static cudaError_t launch_wgrad_tc_mt(const WgradProblem& w, const CUtensorMap& tmY, const CUtensorMap& tmX, int max_kb,
                                      cudaStream_t st) {
  constexpr int BN = 256;
  static int stages_cfg = env_int("MMR_TC_WGRAD_STAGES", MT == 1 ? 2 : 3);
  const int max_stages = MT == 1 ? 4 : 3;
  const int stages = stages_cfg < 1 ? 1 : (stages_cfg > max_stages ? max_stages : stages_cfg);
  const int tiles_per_seg = ((w.M + MT * BM - 1) / (MT * BM)) * ((w.N + BN - 1) / BN);
  const int ctas_per_wave = 148 * (MT == 1 ? 2 : 1);
  // distribute ~one wave of CTAs over the segments in proportion to their reduction length
  long long total_kb = 0;
  for (int s = 0; s < w.segs.n; ++s) total_kb += (w.segs.row0[s + 1] - w.segs.row0[s]) / BK;
  const int units = ctas_per_wave / tiles_per_seg > w.segs.n ? ctas_per_wave / tiles_per_seg : w.segs.n;
  static int prop = env_int("MMR_TC_WGRAD_PROP", 0);
  WgradSplits tab;
  tab.first[0] = 0;
  for (int s = 0; s < w.segs.n; ++s) {
    const int kb = (w.segs.row0[s + 1] - w.segs.row0[s]) / BK;
    int sp = prop ? (int)(((long long)units * kb + total_kb / 2) / (total_kb > 0 ? total_kb : 1))
                  : (ctas_per_wave + tiles_per_seg * w.segs.n - 1) / (tiles_per_seg * w.segs.n);
    if (sp > kb / 4) sp = kb / 4;       // at least 4 k-blocks (256 rows) per chunk
    if (sp < 1) sp = 1;
    tab.first[s + 1] = tab.first[s] + sp;
  }
  for (int s = w.segs.n; s < 6; ++s) tab.first[s + 1] = tab.first[w.segs.n];
  (void)max_kb;
  // opt-in: measured on one box (gpurun_out/r2c21_*): the class alone 1.03 -> 0.93 ms, but the two-stream step 4.65 -> 4.75 ms --
  // balanced chunks keep all 148 SMs busy for the whole launch, and the chain's persistent GEMMs (static tile schedule, one
  // CTA per SM) then wait for every SM; with the static table the short segments' CTAs retire early and free theirs
  static int bal = env_int("MMR_TC_WGRAD_BAL", 0);
  int grid_z = tab.first[w.segs.n];
  if (bal) {        // the kernel derives the chunk table from the device-side row counts (see wgrad_tc_kernel)
    grid_z = ctas_per_wave / tiles_per_seg > w.segs.n ? ctas_per_wave / tiles_per_seg : w.segs.n;
    tab.first[0] = -1;
  }
  auto kern = wgrad_tc_kernel<MT, BN>;
  const int smem = wgrad_smem_bytes<MT, BN>(stages);
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return err;
  dim3 grid((w.N + BN - 1) / BN, (w.M + MT * BM - 1) / (MT * BM), grid_z);
  if (pdl_enabled()) return launch_pdl(kern, grid, dim3(NTHREADS), (size_t)smem, st, tmY, tmX, w, tab, stages);
  kern<<<grid, NTHREADS, smem, st>>>(tmY, tmX, w, tab, stages);
  return cudaGetLastError();
}

static cudaError_t launch_wgrad_tc(const WgradProblem& w, int y_rows_total, int x_rows_total, cudaStream_t st) {
  CUtensorMap tmY, tmX;
  if (!make_tmap(&tmY, w.dY, (uint64_t)w.ldy, (uint64_t)y_rows_total, (uint64_t)w.ldy, 64, BK)) return cudaErrorUnknown;
  if (!make_tmap(&tmX, w.X, (uint64_t)w.ldx, (uint64_t)x_rows_total, (uint64_t)w.ldx, 64, BK)) return cudaErrorUnknown;
  int max_kb = 1;
  for (int s = 0; s < w.segs.n; ++s) {
    int kb = (w.segs.row0[s + 1] - w.segs.row0[s]) / BK;
    if (kb > max_kb) max_kb = kb;
  }
  static int mt_cfg = env_int("MMR_TC_WGRAD_MT", 2);
  if (mt_cfg == 2 && w.M % (2 * BM) == 0) return launch_wgrad_tc_mt<2>(w, tmY, tmX, max_kb, st);
  return launch_wgrad_tc_mt<1>(w, tmY, tmX, max_kb, st);
}

}  // namespace tc
}  // namespace mmr
