// CUDA-core GEMMs: the fp32 parity engine (and a debug engine for bf16 operands).
// Same GemmProblem / epilogue contract as the tcgen05 engine in gemm_tc.cuh.
#pragma once
#include "epilogue.cuh"

namespace mmr {

// C[r, n] = sum_k A[a(r), k] * B[b(seg)+n, k];  64x64xBK tiles, 256 threads, 4x4 per thread.
// Requires K % BK == 0 (BK = 64 when K % 64 == 0, else 16), N % 4 == 0, lda/ldb % 4 == 0.  BK = 64 issues four
// independent 16-byte loads per operand per thread and quarters the number of barrier-separated steps, which
// is what the small (M = batch) pair / trimodal projections are bound by.
// TRANSB: B is a single row-major [K, N] matrix (B(n,k) = B[k*ldb + b_row0 + n]).
// ---- tf32 tensor-core contraction of one staged slab (reduced-precision mode) -------------------------------------------
// The small-M GEMMs of the path (pair / trimodal composition, capsule and projector weight gradients: M = batch or
// M = 32 / 33) are bound by the 16 FMA + 2 LDS.128 per k of the 4x4 register tile, not by bytes: the same slabs
// contracted with mma.sync m16n8k8 tf32 need ~9x fewer instructions.  Operands are k-major in shared memory
// (As[k][m], Bs[k][n], pitch 72 floats: bank = 8 t + g, conflict-free fragment loads); warp w owns m-tile (w & 3) and the four
// n-tiles of column block (w >> 2) of the 64 x 64 tile.
constexpr int SP = 72;
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int KS>
__device__ __forceinline__ void slab_mma_tf32(const float (*As)[SP], const float (*Bs)[SP], int kmax, float (&acc)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    if (ks * 8 >= kmax) break;
    const uint32_t a0 = to_tf32(As[ks * 8 + t][m0 + g]), a1 = to_tf32(As[ks * 8 + t][m0 + g + 8]);
    const uint32_t a2 = to_tf32(As[ks * 8 + t + 4][m0 + g]), a3 = to_tf32(As[ks * 8 + t + 4][m0 + g + 8]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t b0 = to_tf32(Bs[ks * 8 + t][n0 + j * 8 + g]), b1 = to_tf32(Bs[ks * 8 + t + 4][n0 + j * 8 + g]);
      mma_tf32(acc[j], a0, a1, a2, a3, b0, b1);
    }
  }
}
// C fragments -> the 4 x 4 register tile of the CUDA-core epilogue (through Cs[m][n], aliasing the A slab)
__device__ __forceinline__ void frags_to_tile(float (*Cs)[SP], const float (&frag)[4][4], float (&acc)[4][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    Cs[m0 + g][n0 + j * 8 + 2 * t] = frag[j][0]; Cs[m0 + g][n0 + j * 8 + 2 * t + 1] = frag[j][1];
    Cs[m0 + g + 8][n0 + j * 8 + 2 * t] = frag[j][2]; Cs[m0 + g + 8][n0 + j * 8 + 2 * t + 1] = frag[j][3];
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(&Cs[ty * 4 + i][tx * 4]);
    acc[i][0] = v.x; acc[i][1] = v.y; acc[i][2] = v.z; acc[i][3] = v.w;
  }
}

template <class TA, class TB, int OP, class CT, bool TRANSB, int BK>
__device__ __forceinline__ void gemm_simt_body(const GemmProblem& g, const EpiParams& e, int bx, int by) {
  constexpr int NL = BK / 16;     // 16-wide k slabs per step
  __shared__ __align__(16) float As[BK < 64 ? 64 : BK][SP];     // >= 64 rows: reused as the C staging tile of the tf32 path
  __shared__ __align__(16) float Bs[BK][SP];
  const int t = threadIdx.x;
  const int m0 = by * 64;
  const int n0 = bx * 64;
  const int seg = seg_of_row(g.segs, m0);
  const int local0 = m0 - g.segs.row0[seg];
  const int rows_valid = seg_rows(g.segs, seg) - local0;  // may be <= 0 for pure padding tiles
  const TA* A = reinterpret_cast<const TA*>(g.A);
  const TB* B = reinterpret_cast<const TB*>(g.B);
  const int lm = t >> 2, lk = (t & 3) * 4;
  const bool a_ok = lm < rows_valid;
  const bool b_ok = (n0 + lm) < g.N;
  const TA* ap = A + (size_t)(g.a_row0[seg] + local0 + lm) * g.lda + lk;
  const TB* bp = B + (size_t)(g.b_row0[seg] + n0 + lm) * g.ldb + lk;
  const int tk = t >> 4, tn = (t & 15) * 4;             // TRANSB load mapping
  const bool bt_ok = (n0 + tn) < g.N;
  const TB* btp = B + (size_t)tk * g.ldb + g.b_row0[seg] + n0 + tn;
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float frag[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) frag[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    float4 av[NL], bv[NL];
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      av[u] = a_ok ? Vec4<TA>::ld(ap + k0 + 16 * u) : zero4;
      if (TRANSB) bv[u] = bt_ok ? Vec4<TB>::ld(btp + (size_t)(k0 + 16 * u) * g.ldb) : zero4;
      else bv[u] = b_ok ? Vec4<TB>::ld(bp + k0 + 16 * u) : zero4;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < NL; ++u) {
      const int kk = 16 * u + lk;
      As[kk + 0][lm] = av[u].x; As[kk + 1][lm] = av[u].y; As[kk + 2][lm] = av[u].z; As[kk + 3][lm] = av[u].w;
      if (TRANSB) {
        *reinterpret_cast<float4*>(&Bs[16 * u + tk][tn]) = bv[u];
      } else {
        Bs[kk + 0][lm] = bv[u].x; Bs[kk + 1][lm] = bv[u].y; Bs[kk + 2][lm] = bv[u].z; Bs[kk + 3][lm] = bv[u].w;
      }
    }
    __syncthreads();
    if (g.tf32) {
      slab_mma_tf32<BK / 8>(As, Bs, BK, frag);
      continue;
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
  if (g.tf32) frags_to_tile(As, frag, acc);
  const int n = n0 + tx * 4;
  if (n < g.N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lr = ty * 4 + i;
      const int crow = m0 + lr;
      if (crow < g.segs.row0[seg + 1])
        epi_apply<OP, CT>(e, crow, lr < rows_valid, g.b_row0[seg] + n, n,
                          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

template <class TA, class TB, int OP, class CT, bool TRANSB, int BK>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmProblem g, EpiParams e) {
  gemm_simt_body<TA, TB, OP, CT, TRANSB, BK>(g, e, blockIdx.x, blockIdx.y);
}

// Up to 3 independent small GEMMs of identical shape in one launch (blockIdx.z = problem): the pair
// projections of the trimodal composition (mult_model.py:174-177) and their data gradients.
struct MultiGemm {
  int n;
  GemmProblem g[3];
  EpiParams e[3];
};
template <class TA, class TB, int OP, class CT, bool TRANSB, int BK>
__global__ void __launch_bounds__(256) gemm_simt_multi_kernel(MultiGemm m) {
  gemm_simt_body<TA, TB, OP, CT, TRANSB, BK>(m.g[blockIdx.z], m.e[blockIdx.z], blockIdx.x, blockIdx.y);
}
template <class TA, class TB, int OP, class CT, bool TRANSB>
static void launch_gemm_simt_multi(const MultiGemm& m, cudaStream_t st) {
  const GemmProblem& g = m.g[0];
  dim3 grid((g.N + 63) / 64, (g.segs.row0[g.segs.n] + 63) / 64, m.n);
  if (g.K % 64 == 0) gemm_simt_multi_kernel<TA, TB, OP, CT, TRANSB, 64><<<grid, 256, 0, st>>>(m);
  else gemm_simt_multi_kernel<TA, TB, OP, CT, TRANSB, 16><<<grid, 256, 0, st>>>(m);
}

template <class TA, class TB, int OP, class CT, bool TRANSB = false>
static void launch_gemm_simt(const GemmProblem& g, const EpiParams& e, cudaStream_t st) {
  const int total_rows = g.segs.row0[g.segs.n];
  dim3 grid((g.N + 63) / 64, (total_rows + 63) / 64);
  if (g.K % 64 == 0) gemm_simt_kernel<TA, TB, OP, CT, TRANSB, 64><<<grid, 256, 0, st>>>(g, e);
  else gemm_simt_kernel<TA, TB, OP, CT, TRANSB, 16><<<grid, 256, 0, st>>>(g, e);
}

// Weight-gradient tile: out[m, n] += sum_{r in [r_begin, r_end)} dY[r, m] * X[r, n]; fully guarded
// scalar loads so that odd M/N/strides (projector 33x256, capsule votes) work.
template <class TY, class TX>
__device__ __forceinline__ void wgrad_tile(const TY* dY, int ldy, const TX* X, int ldx, int r_begin, int r_end,
                                           int M, int N, int m0, int n0, float* out, int ldo, int tf32 = 0) {
  constexpr int RB = 64;            // reduction rows staged per barrier-separated step (16 loads in flight per thread)
  __shared__ __align__(16) float Ys[RB][SP];
  __shared__ __align__(16) float Xs[RB][SP];
  float frag[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) frag[i][j] = 0.f;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int c = t & 63, k0 = t >> 6;          // staging: thread covers column c, rows k0, k0+4, ...
  const bool y_ok = m0 + c < M, x_ok = n0 + c < N;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += RB) {
    float yv[RB / 4], xv[RB / 4];
#pragma unroll
    for (int i = 0; i < RB / 4; ++i) {
      const int r = r0 + k0 + 4 * i;
      yv[i] = (r < r_end && y_ok) ? to_f<TY>(dY[(size_t)r * ldy + m0 + c]) : 0.f;
      xv[i] = (r < r_end && x_ok) ? to_f<TX>(X[(size_t)r * ldx + n0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RB / 4; ++i) { Ys[k0 + 4 * i][c] = yv[i]; Xs[k0 + 4 * i][c] = xv[i]; }
    __syncthreads();
    const int kmax = min(RB, r_end - r0);
    if (tf32) {      // rows beyond r_end were staged as zeros
      slab_mma_tf32<RB / 8>(Ys, Xs, kmax, frag);
      continue;
    }
#pragma unroll 8
    for (int k = 0; k < kmax; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = Ys[k][ty * 4 + i]; b[i] = Xs[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  if (out == nullptr) return;
  if (tf32) frags_to_tile(Ys, frag, acc);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) atomicAdd(out + (size_t)m * ldo + n, acc[i][j]);
    }
  }
}

// grid.z = nseg * splits
template <class TY, class TX>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradProblem w, int splits) {
  const int seg = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int rows = seg_rows(w.segs, seg);
  const int chunk = (((rows + splits - 1) / splits) + 15) / 16 * 16;
  const int r_begin = sp * chunk;
  const int r_end = min(rows, r_begin + chunk);
  const TY* dY = reinterpret_cast<const TY*>(w.dY) + (size_t)w.segs.row0[seg] * w.ldy;
  const TX* X = reinterpret_cast<const TX*>(w.X) + (size_t)w.x_row0[seg] * w.ldx;
  wgrad_tile<TY, TX>(dY, w.ldy, X, w.ldx, r_begin, r_end, w.M, w.N, blockIdx.y * 64, blockIdx.x * 64,
                     w.out[seg], w.ldo, w.tf32);
}

// Batched fp32 weight gradients over one shared row range (capsule votes / projector): batch i has
// its own operand bases.  grid.z = nbatch * splits.
struct WgradBatch {
  int nbatch, rows, M, N, ldy, ldx, ldo, tf32;
  const float* dY[MMR_ROUTES];
  const float* X[MMR_ROUTES];
  float* out[MMR_ROUTES];
};
__global__ void __launch_bounds__(256) wgrad_batched_kernel(WgradBatch w, int splits) {
  const int bi = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int chunk = (((w.rows + splits - 1) / splits) + 15) / 16 * 16;
  const int r_begin = sp * chunk;
  const int r_end = min(w.rows, r_begin + chunk);
  wgrad_tile<float, float>(w.dY[bi], w.ldy, w.X[bi], w.ldx, r_begin, r_end, w.M, w.N, blockIdx.y * 64,
                           blockIdx.x * 64, w.out[bi], w.ldo, w.tf32);
}
static void launch_wgrad_batched(const WgradBatch& w, cudaStream_t st) {
  const int tiles = ((w.M + 63) / 64) * ((w.N + 63) / 64) * w.nbatch;
  int splits = (148 * 4 + tiles - 1) / tiles;
  const int max_splits = (w.rows + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  dim3 grid((w.N + 63) / 64, (w.M + 63) / 64, w.nbatch * splits);
  wgrad_batched_kernel<<<grid, 256, 0, st>>>(w, splits);
}

template <class TY, class TX>
static void launch_wgrad_simt(const WgradProblem& w, cudaStream_t st) {
  int max_rows = 1;
  for (int s = 0; s < w.segs.n; ++s) max_rows = max_rows > w.segs.rows[s] ? max_rows : w.segs.rows[s];
  const int tiles = ((w.M + 63) / 64) * ((w.N + 63) / 64) * w.segs.n;
  int splits = (148 * 4 + tiles - 1) / tiles;
  const int max_splits = (max_rows + 63) / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  dim3 grid((w.N + 63) / 64, (w.M + 63) / 64, w.segs.n * splits);
  wgrad_simt_kernel<TY, TX><<<grid, 256, 0, st>>>(w, splits);
}

}  // namespace mmr
