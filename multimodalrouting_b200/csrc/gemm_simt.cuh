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
template <class TA, class TB, int OP, class CT, bool TRANSB, int BK>
__device__ __forceinline__ void gemm_simt_body(const GemmProblem& g, const EpiParams& e, int bx, int by) {
  constexpr int NL = BK / 16;     // 16-wide k slabs per step
  __shared__ __align__(16) float As[BK][68];
  __shared__ __align__(16) float Bs[BK][68];
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
                                           int M, int N, int m0, int n0, float* out, int ldo) {
  constexpr int RB = 64;            // reduction rows staged per barrier-separated step (16 loads in flight per thread)
  __shared__ float Ys[RB][65];
  __shared__ float Xs[RB][65];
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
                     w.out[seg], w.ldo);
}

// Batched fp32 weight gradients over one shared row range (capsule votes / projector): batch i has
// its own operand bases.  grid.z = nbatch * splits.
struct WgradBatch {
  int nbatch, rows, M, N, ldy, ldx, ldo;
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
                           blockIdx.x * 64, w.out[bi], w.ldo);
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
