// Element-wise GEMM epilogues shared by the SIMT (fp32 parity) and tcgen05 (bf16) engines.
#pragma once
#include "mmr_common.cuh"

namespace mmr {

enum EpiOp {
  EPI_BIAS = 0,             // out<CT>  = acc + bias                      (Q, K/V projections)
  EPI_BIAS_RELU = 1,        // out<CT>  = relu(acc + bias)                (fc1)
  EPI_BIAS_RESID_MASK = 2,  // out<f32> = (resid + acc + bias) * rowmask  (out_proj, fc2 + residual)
  EPI_RELUMASK = 3,         // out<CT>  = acc * (aux > 0)                 (d fc1-output)
  EPI_MASK = 4,             // out<CT>  = acc * rowmask (rowmask may be null)
  EPI_STORE_F32 = 5,        // out<f32> = acc
  EPI_BIAS_F32 = 6,         // out<f32> = acc + bias (bias may be null)
};

// One grouped GEMM: C rows are split into segments (directions); each segment multiplies its A
// rows with its own block of the stacked B matrix:  C[r, n] = sum_k A[a(r), k] * B[b_row0[seg]+n, k]
struct GemmProblem {
  Segs segs;        // C row space
  int a_row0[6];    // A row of the first C row of each segment
  int b_row0[6];    // first row of the segment's block in the stacked B
  int N, K;
  const void* A;
  int lda;
  const void* B;
  int ldb;
  int tf32;         // CUDA-core engine only: 1 = the staged slabs are contracted with mma.sync tf32 (reduced-precision mode;
                    // the fp32 parity mode keeps the FMA loop)
};

struct EpiParams {
  const float* bias;     // indexed like stacked B rows (b_row0[seg] + n); may be null
  const float* resid;    // fp32 [rows, ldr], C row space
  int ldr;
  const float* rowmask;  // fp32 [rows], C row space; may be null
  const void* aux;       // CT [rows, ldaux]
  int ldaux;
  void* out;
  int ldo;
};

// acc: 4 consecutive columns n..n+3 of C row `crow`.  `valid` is false for the zero-padded tail
// rows of a segment, which are written as zeros so later reductions over rows stay exact.
template <int OP, class CT>
__device__ __forceinline__ void epi_apply(const EpiParams& p, int crow, bool valid, int bcol, int n,
                                          float4 acc) {
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    if (OP == EPI_BIAS || OP == EPI_BIAS_RELU || OP == EPI_BIAS_RESID_MASK || OP == EPI_BIAS_F32) {
      if (p.bias != nullptr) {
        float4 b = *reinterpret_cast<const float4*>(p.bias + bcol);
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
      }
    }
    if (OP == EPI_BIAS_RELU) {
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f);
      acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    }
    if (OP == EPI_BIAS_RESID_MASK) {
      float4 r = *reinterpret_cast<const float4*>(p.resid + (size_t)crow * p.ldr + n);
      float m = p.rowmask ? p.rowmask[crow] : 1.f;
      acc.x = (r.x + acc.x) * m; acc.y = (r.y + acc.y) * m;
      acc.z = (r.z + acc.z) * m; acc.w = (r.w + acc.w) * m;
    }
    if (OP == EPI_RELUMASK) {
      float4 f = Vec4<CT>::ld(reinterpret_cast<const CT*>(p.aux) + (size_t)crow * p.ldaux + n);
      acc.x = f.x > 0.f ? acc.x : 0.f; acc.y = f.y > 0.f ? acc.y : 0.f;
      acc.z = f.z > 0.f ? acc.z : 0.f; acc.w = f.w > 0.f ? acc.w : 0.f;
    }
    if (OP == EPI_MASK) {
      float m = p.rowmask ? p.rowmask[crow] : 1.f;
      acc.x *= m; acc.y *= m; acc.z *= m; acc.w *= m;
    }
    o = acc;
  }
  if (OP == EPI_BIAS_RESID_MASK || OP == EPI_STORE_F32 || OP == EPI_BIAS_F32) {
    Vec4<float>::st(reinterpret_cast<float*>(p.out) + (size_t)crow * p.ldo + n, o);
  } else {
    Vec4<CT>::st(reinterpret_cast<CT*>(p.out) + (size_t)crow * p.ldo + n, o);
  }
}

// Weight-gradient GEMM: out[seg][m, n] += scale * sum_{r in seg} dY[r, m] * X[x(r), n]
struct WgradProblem {
  Segs segs;          // reduction-row segments in dY's row space
  int x_row0[6];      // X row of the first dY row of each segment
  const void* dY;
  int ldy;
  const void* X;
  int ldx;
  int M, N;
  float* out[6];      // per segment fp32 [M, ldo] accumulators (zero-initialised by the caller)
  int ldo;
  float* dbias[6];    // optional per segment fp32 [M]: += column sums of dY (bias gradient), fused into the
  int colsum;         // tcgen05 weight-gradient kernel when `colsum` is set (launch-uniform)
  int tf32;           // CUDA-core engine only: contract the staged slabs with mma.sync tf32 (reduced-precision mode)
};

}  // namespace mmr
