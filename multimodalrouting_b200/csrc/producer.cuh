// Route-input producers that sit between the modality encoders and the hot path:
//   BioClinicalBERT chunk projection  proj = Sequential(LayerNorm(768), Linear(768 -> 256, bias=False))
//                                     MIMIC-IV/MortModel/Paired_Cross_Attention/encoders.py:289-293, 472-475
//   CXR token projection              token_proj = Linear(512 -> 256, bias=False)          encoders.py:620, 747-749
// Row kernels of this file: (optional) LayerNorm over a D-wide row + cast to the GEMM operand type, and its backward.
// The Linear itself runs on the engines of gemm_tc.cuh / gemm_simt.cuh (api.cu: mmr_producer_proj_fwd / bwd).
#pragma once
#include "mmr_common.cuh"

namespace mmr {

constexpr int PR_MAXV = 8;     // float4 groups per lane: D <= 32 * 4 * 8 = 1024, D % 128 == 0

template <class XT> __device__ __forceinline__ float4 pr_ld4(const void* base, size_t idx);
template <> __device__ __forceinline__ float4 pr_ld4<float>(const void* base, size_t idx) {
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
}
template <> __device__ __forceinline__ float4 pr_ld4<bf16>(const void* base, size_t idx) {
  return Vec4<bf16>::ld(reinterpret_cast<const bf16*>(base) + idx);
}

struct ProjRowArgs {
  const void* x; long long rows; int rows_pad; int D; int has_ln;
  const float* gamma; const float* beta;
  void* out;           // OT [rows_pad, D]: LayerNorm(x) * gamma + beta (or x), pad rows zero
  float* stat;         // [rows, 2] mean, rstd (has_ln)
  // backward
  const float* dh;     // fp32 [rows_pad, D]: gradient wrt the LayerNorm output
  float* dx;           // fp32 [rows, D] (may be null)
  float* dgamma; float* dbeta;   // fp32 [D] accumulators (may be null)
};

// one warp per row; lane owns the float4 groups lane, lane + 32, ...
template <class XT, class OT>
__global__ void __launch_bounds__(256) proj_rows_fwd_kernel(ProjRowArgs a) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= a.rows_pad) return;
  const int nv = a.D / 128;
  OT* orow = reinterpret_cast<OT*>(a.out) + (size_t)r * a.D;
  if (r >= a.rows) {
    for (int i = 0; i < nv; ++i) Vec4<OT>::st(orow + 4 * (lane + 32 * i), make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  float4 v[PR_MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PR_MAXV; ++i)
    if (i < nv) {
      v[i] = pr_ld4<XT>(a.x, (size_t)r * a.D + 4 * (lane + 32 * i));
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  if (a.has_ln) {
    const float mean = warp_sum(s) / (float)a.D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PR_MAXV; ++i)
      if (i < nv) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        q += dx * dx + dy * dy + dz * dz + dw * dw;
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)a.D + LN_EPS);
#pragma unroll
    for (int i = 0; i < PR_MAXV; ++i)
      if (i < nv) {
        const int c = 4 * (lane + 32 * i);
        const float4 g = *reinterpret_cast<const float4*>(a.gamma + c), b = *reinterpret_cast<const float4*>(a.beta + c);
        v[i] = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                           (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      }
    if (lane == 0) { a.stat[2 * r] = mean; a.stat[2 * r + 1] = rstd; }
  }
#pragma unroll
  for (int i = 0; i < PR_MAXV; ++i)
    if (i < nv) Vec4<OT>::st(orow + 4 * (lane + 32 * i), v[i]);
}

// dx = rstd * (dh*gamma - mean(dh*gamma) - xhat * mean(dh*gamma*xhat)); dgamma += dh * xhat; dbeta += dh.
// Blocks walk row stripes and keep the column sums in registers; one atomicAdd per (block, column).
template <class XT>
__global__ void __launch_bounds__(256) proj_rows_bwd_kernel(ProjRowArgs a, int rows_per_block) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = a.D / 128;
  float4 dg[PR_MAXV], db[PR_MAXV];
#pragma unroll
  for (int i = 0; i < PR_MAXV; ++i) { dg[i] = make_float4(0.f, 0.f, 0.f, 0.f); db[i] = dg[i]; }
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(a.rows, r_begin + rows_per_block);
  for (long long r = r_begin + warp; r < r_end; r += 8) {
    const float mean = a.stat[2 * r], rstd = a.stat[2 * r + 1];
    float4 xh[PR_MAXV], gy[PR_MAXV];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < PR_MAXV; ++i)
      if (i < nv) {
        const int c = 4 * (lane + 32 * i);
        const float4 x = pr_ld4<XT>(a.x, (size_t)r * a.D + c);
        const float4 d = *reinterpret_cast<const float4*>(a.dh + (size_t)r * a.D + c);
        const float4 g = *reinterpret_cast<const float4*>(a.gamma + c);
        xh[i] = make_float4((x.x - mean) * rstd, (x.y - mean) * rstd, (x.z - mean) * rstd, (x.w - mean) * rstd);
        gy[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        m1 += gy[i].x + gy[i].y + gy[i].z + gy[i].w;
        m2 += gy[i].x * xh[i].x + gy[i].y * xh[i].y + gy[i].z * xh[i].z + gy[i].w * xh[i].w;
        dg[i].x += d.x * xh[i].x; dg[i].y += d.y * xh[i].y; dg[i].z += d.z * xh[i].z; dg[i].w += d.w * xh[i].w;
        db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
      }
    m1 = warp_sum(m1) / (float)a.D;
    m2 = warp_sum(m2) / (float)a.D;
    if (a.dx)
#pragma unroll
      for (int i = 0; i < PR_MAXV; ++i)
        if (i < nv) {
          const int c = 4 * (lane + 32 * i);
          *reinterpret_cast<float4*>(a.dx + (size_t)r * a.D + c) =
              make_float4(rstd * (gy[i].x - m1 - xh[i].x * m2), rstd * (gy[i].y - m1 - xh[i].y * m2),
                          rstd * (gy[i].z - m1 - xh[i].z * m2), rstd * (gy[i].w - m1 - xh[i].w * m2));
        }
  }
  // the 8 warps of the block hold partial column sums for the same columns: combine through shared memory
  __shared__ float red[8][128];
  for (int i = 0; i < nv; ++i) {
    for (int pass = 0; pass < 2; ++pass) {
      float* dst = pass == 0 ? a.dgamma : a.dbeta;
      if (dst == nullptr) continue;
      const float4 v = pass == 0 ? dg[i] : db[i];
      __syncthreads();
      *reinterpret_cast<float4*>(&red[warp][4 * lane]) = v;
      __syncthreads();
      if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        // column of element j of group (lane', i): 4 * (lane' + 32 i) + j  with 4 lane' + j = threadIdx.x
        atomicAdd(dst + 128 * i + threadIdx.x, s);
      }
    }
  }
}

// fp32 [rows, D] -> OT [rows_pad, D] (pad rows zero): the bf16 operand copy of an incoming gradient
template <class OT>
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* src, long long rows, int rows_pad, int D, void* dst) {
  const size_t n4 = (size_t)rows_pad * D / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const size_t r = i * 4 / D;
    const float4 v = r < (size_t)rows ? *reinterpret_cast<const float4*>(src + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    Vec4<OT>::st(reinterpret_cast<OT*>(dst) + i * 4, v);
  }
}

}  // namespace mmr
