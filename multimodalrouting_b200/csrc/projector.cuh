// Standalone RoutePrimaryProjector (routing_and_heads.py:101-121): 10 independent Linear(256 -> 33); pose = out[:32],
// act = sigmoid(out[32]).  The hot path never calls it (forward_capsule_from_route_dict runs the projector inside the
// routing kernel); this is the module's own forward() for callers that use the projector alone.  fp32 throughout.
#pragma once
#include "mmr_common.cuh"

namespace mmr {

constexpr int PJ_PB = 8;   // patients per CTA

struct ProjectorArgs {
  const float* w[MMR_ROUTES];    // [33,256]
  const float* b[MMR_ROUTES];    // [33]
  const float* embs; long long rs, bs; int B;
  float* poses; float* acts;             // fwd out: [B,10,32], [B,10]
  const float* d_poses; const float* d_acts;   // bwd in (either may be null)
  float* dpc;                            // bwd: [B,10,33] gradient wrt the projector outputs (operand of dW / db)
  float* d_embs;                         // bwd out, same strides as embs (may be null)
};

// grid (ceil(B / PJ_PB), 10), 256 threads: the tile's 8 embedding rows are staged once, each warp owns output rows
// j = warp, warp + 8, ... and keeps its weight row in registers across the 8 patients
__global__ void __launch_bounds__(256) projector_fwd_kernel(ProjectorArgs a) {
  __shared__ float e[PJ_PB][256];
  const int r = blockIdx.y, b0 = blockIdx.x * PJ_PB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = min(PJ_PB, a.B - b0);
  for (int i = tid; i < PJ_PB * 256; i += 256) {
    const int p = i >> 8, c = i & 255;
    e[p][c] = p < np ? a.embs[(size_t)r * a.rs + (size_t)(b0 + p) * a.bs + c] : 0.f;
  }
  __syncthreads();
  for (int j = warp; j < 33; j += 8) {
    float wv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[i] = a.w[r][(size_t)j * 256 + lane + 32 * i];
    const float bias = a.b[r][j];
    for (int p = 0; p < np; ++p) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(wv[i], e[p][lane + 32 * i], acc);
      acc = warp_sum(acc) + bias;
      if (lane == 0) {
        if (j < 32) a.poses[((size_t)(b0 + p) * 10 + r) * 32 + j] = acc;
        else a.acts[(size_t)(b0 + p) * 10 + r] = 1.0f / (1.0f + expf(-acc));
      }
    }
  }
}

// same grid: dpc[b,r,j] = d_poses[b,r,j] (j < 32) | d_acts[b,r] * a (1 - a) (j = 32), a recomputed from the embeddings;
// d_embs[r,b,c] = sum_j dpc[b,r,j] W_r[j,c]  (thread = column)
__global__ void __launch_bounds__(256) projector_bwd_kernel(ProjectorArgs a) {
  __shared__ float e[PJ_PB][256];
  __shared__ float g[PJ_PB][36];
  const int r = blockIdx.y, b0 = blockIdx.x * PJ_PB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int np = min(PJ_PB, a.B - b0);
  for (int i = tid; i < PJ_PB * 256; i += 256) {
    const int p = i >> 8, c = i & 255;
    e[p][c] = p < np ? a.embs[(size_t)r * a.rs + (size_t)(b0 + p) * a.bs + c] : 0.f;
  }
  for (int i = tid; i < PJ_PB * 32; i += 256) {
    const int p = i >> 5, j = i & 31;
    g[p][j] = (p < np && a.d_poses) ? a.d_poses[((size_t)(b0 + p) * 10 + r) * 32 + j] : 0.f;
  }
  __syncthreads();
  if (warp < np) {     // activation logit of patient `warp`, then d logit = d act * a (1 - a)
    const int p = warp;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(a.w[r][(size_t)32 * 256 + lane + 32 * i], e[p][lane + 32 * i], acc);
    acc = warp_sum(acc) + a.b[r][32];
    if (lane == 0) {
      const float s = 1.0f / (1.0f + expf(-acc));
      g[p][32] = a.d_acts ? a.d_acts[(size_t)(b0 + p) * 10 + r] * s * (1.0f - s) : 0.f;
    }
  }
  __syncthreads();
  for (int i = tid; i < np * 33; i += 256) {
    const int p = i / 33, j = i % 33;
    a.dpc[((size_t)(b0 + p) * 10 + r) * 33 + j] = g[p][j];
  }
  if (a.d_embs) {
    float acc[PJ_PB];
#pragma unroll
    for (int p = 0; p < PJ_PB; ++p) acc[p] = 0.f;
    for (int j = 0; j < 33; ++j) {
      const float wv = a.w[r][(size_t)j * 256 + tid];
#pragma unroll
      for (int p = 0; p < PJ_PB; ++p) acc[p] = fmaf(g[p][j], wv, acc[p]);
    }
    for (int p = 0; p < np; ++p) a.d_embs[(size_t)r * a.rs + (size_t)(b0 + p) * a.bs + tid] = acc[p];
  }
}

}  // namespace mmr
