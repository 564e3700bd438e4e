// Capsule-style routing + heads: one persistent CTA walks patients; all per-patient routing state
// (10 route embeddings, primary capsules, 10 x K x 64 votes, per-iteration decision poses and
// routing coefficients) lives in shared memory across all agreement iterations.
//   RoutePrimaryProjector.forward         routing_and_heads.py:111-121
//   forward_capsule_from_route_dict       routing_and_heads.py:314-352 (mask / temperature / clamp)
//   CapsuleMortalityHead.forward          Mort :194-268 / Pheno :194-272
//   CapsuleFC.forward                     capsule_layers.py:75-117
#pragma once
#include "mmr_common.cuh"

namespace mmr {

constexpr int RT_THREADS = 256;
constexpr int RT_MAXIT = 4;

struct RoutingArgs {
  mmr_routing_dims d;
  mmr_routing_params p;
  const float* route_embs; const float* poses_in; const float* acts_in;
  const float* acts_override; const float* route_mask;
  float* logits; float* alpha; float* R; float* poses_out; float* acts_out;
  // backward
  const float* d_logits; const float* d_R;
  float* d_route_embs; float* d_poses; float* d_acts;
  float* du;        // [B,10,K*64] gradient wrt votes
  float* poses_m;   // [B,10,32]  masked primary poses (operand of the vote-weight gradient)
  float* dpc;       // [B,10,33]  gradient wrt projector outputs
  float* dG;        // [K,32]     accumulated gradient wrt G = embedding @ pose_to_mc
  float* dbias;     // [K] or null
};

// shared-memory carve (floats)
struct RtSmem {
  float* e;      // [10][256]
  float* pose;   // [10][32]  masked poses
  float* zl;     // [10] projector activation logits
  float* a0;     // [10] sigmoid / override
  float* a2;     // [10] after temperature (pre-clamp)
  float* a3;     // [10] prior after clamp
  float* alpha;  // [10] prim_act * mask
  float* act;    // [10] activation used inside routing
  float* rm;     // [10]
  float* G;      // [K][32]
  float* u;      // [10][K*64]
  float* v;      // [nit][K*64]
  float* q;      // [nit][10*K]
  float* Rn;     // [10*K]
  float* dp;     // [K][32] decision poses d_bkp
  // backward only
  float* dv;     // [nit][K*64]
  float* ds;     // [nit][10*K]
  float* ddp;    // [K][32]
  float* dpose;  // [10][32]
  float* misc;   // [64]
};

__host__ __device__ inline size_t rt_smem_floats(int K, int nit, bool bwd) {
  size_t n = 10 * 256 + 10 * 32 + 7 * 16 + K * 32 + 10 * K * 64 + (size_t)nit * K * 64 + (size_t)nit * 10 * K + 10 * K + K * 32;
  if (bwd) n += (size_t)nit * K * 64 + (size_t)nit * 10 * K + K * 32 + 10 * 32 + 64;
  return n + 64;
}

__device__ inline RtSmem rt_carve(float* base, int K, int nit, bool bwd) {
  RtSmem s;
  float* p = base;
  s.e = p; p += 10 * 256;
  s.pose = p; p += 10 * 32;
  s.zl = p; p += 16; s.a0 = p; p += 16; s.a2 = p; p += 16; s.a3 = p; p += 16;
  s.alpha = p; p += 16; s.act = p; p += 16; s.rm = p; p += 16;
  s.G = p; p += K * 32;
  s.u = p; p += 10 * K * 64;
  s.v = p; p += nit * K * 64;
  s.q = p; p += nit * 10 * K;
  s.Rn = p; p += 10 * K;
  s.dp = p; p += K * 32;
  if (bwd) {
    s.dv = p; p += nit * K * 64;
    s.ds = p; p += nit * 10 * K;
    s.ddp = p; p += K * 32;
    s.dpose = p; p += 10 * 32;
    s.misc = p; p += 64;
  } else {
    s.dv = s.ds = s.ddp = s.dpose = s.misc = nullptr;
  }
  return s;
}

// G[k][p] = sum_m pose_to_mc[m][p] * embedding[k][m]   (logits = sum_p d[k][p]*G[k][p] + bias[k])
__device__ inline void rt_build_G(const RoutingArgs& a, const RtSmem& s) {
  const int K = a.d.K;
  for (int i = threadIdx.x; i < K * 32; i += RT_THREADS) {
    const int k = i >> 5, p = i & 31;
    float acc = 0.f;
    for (int m = 0; m < MC; ++m) acc = fmaf(a.p.pose_to_mc[m * PC + p], a.p.embedding[k * MC + m], acc);
    s.G[i] = acc;
  }
}

// Forward for patient b entirely in shared memory.  Leaves every intermediate in `s`.
__device__ inline void rt_forward(const RoutingArgs& a, const RtSmem& s, int b) {
  const int K = a.d.K, KD = K * 64, nit = a.d.num_routing;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_mask = a.route_mask != nullptr;
  __syncthreads();
  if (tid < 10) s.rm[tid] = has_mask ? a.route_mask[(size_t)b * 10 + tid] : 1.f;
  if (!a.d.from_poses) {
    for (int i = tid; i < 10 * 64; i += RT_THREADS) {
      const int r = i >> 6, c = (i & 63) * 4;
      *reinterpret_cast<float4*>(s.e + r * 256 + c) = *reinterpret_cast<const float4*>(
          a.route_embs + (size_t)r * a.d.emb_route_stride + (size_t)b * a.d.emb_batch_stride + c);
    }
    __syncthreads();
    // projector: 10 x 33 dot products of length 256, one warp each
    for (int o = warp; o < 330; o += RT_THREADS / 32) {
      const int r = o / 33, j = o % 33;
      const float* w = a.p.proj_w[r] + (size_t)j * 256;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(w[lane + 32 * i], s.e[r * 256 + lane + 32 * i], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        acc += a.p.proj_b[r][j];
        if (j < 32) s.pose[r * 32 + j] = acc; else s.zl[r] = acc;
      }
    }
    __syncthreads();
    if (tid < 10) {
      const int r = tid;
      const float rm = s.rm[r];
      float a0 = 1.0f / (1.0f + expf(-s.zl[r]));
      if (a.acts_out) a.acts_out[(size_t)b * 10 + r] = a0;
      if (a.acts_override) a0 = a.acts_override[(size_t)b * 10 + r];
      s.a0[r] = a0;
      float x = has_mask ? a0 * rm : a0;
      const bool keep = has_mask ? (rm != 0.f) : true;
      if (a.d.act_temperature != 1.0f && has_mask && keep) {
        const float xc = fminf(fmaxf(x, 1e-6f), 1.0f - 1e-6f);
        const float lg = (logf(xc) - log1pf(-xc)) / a.d.act_temperature;
        x = 1.0f / (1.0f + expf(-lg));
      }
      s.a2[r] = x;
      if (keep) x = fminf(fmaxf(x, a.d.prior_floor), a.d.prior_ceiling);
      s.a3[r] = x;
      s.alpha[r] = has_mask ? x * rm : x;
    }
    if (a.poses_out)
      for (int i = tid; i < 320; i += RT_THREADS) a.poses_out[(size_t)b * 320 + i] = s.pose[i];
    __syncthreads();
  } else {
    for (int i = tid; i < 320; i += RT_THREADS) s.pose[i] = a.poses_in[(size_t)b * 320 + i];
    if (tid < 10) {
      const float x = a.acts_in[(size_t)b * 10 + tid];
      s.a0[tid] = s.a2[tid] = s.a3[tid] = x;
      s.alpha[tid] = has_mask ? x * s.rm[tid] : x;
    }
    __syncthreads();
  }
  // mask poses, pick the routing activation
  for (int i = tid; i < 320; i += RT_THREADS) s.pose[i] *= s.rm[i >> 5];
  if (tid < 10) s.act[tid] = (a.d.variant == MMR_VARIANT_PHENO) ? s.alpha[tid] : s.rm[tid];
  __syncthreads();
  // votes u[r][c] = sum_a pose[r][a] * w[r][a][c],  c = k*64 + d
  for (int c = tid; c < KD; c += RT_THREADS) {
    for (int r = 0; r < 10; ++r) {
      const float* w = a.p.caps_w + (size_t)r * 32 * KD + c;
      float acc = 0.f;
#pragma unroll 8
      for (int aa = 0; aa < 32; ++aa) acc = fmaf(s.pose[r * 32 + aa], w[(size_t)aa * KD], acc);
      s.u[r * KD + c] = acc;
    }
  }
  __syncthreads();
  const float scale = 0.125f;   // 1/sqrt(64)
  const float invK = 1.0f / (float)K;
  for (int it = 0; it < nit; ++it) {
    float* q = s.q + it * 10 * K;
    if (it == 0) {
      for (int i = tid; i < 10 * K; i += RT_THREADS) q[i] = invK;
    } else {
      const float* vp = s.v + (it - 1) * KD;
      for (int o = warp; o < 10 * K; o += RT_THREADS / 32) {
        const int r = o / K, k = o % K;
        const float* uu = s.u + r * KD + k * 64;
        float acc = uu[lane] * vp[k * 64 + lane] + uu[lane + 32] * vp[k * 64 + lane + 32];
        acc = warp_sum(acc);
        if (lane == 0) q[o] = acc * scale;
      }
      __syncthreads();
      if (tid < 10) {
        float* qr = q + tid * K;
        float mx = -INFINITY;
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, qr[k]);
        float sum = 0.f;
        for (int k = 0; k < K; ++k) { const float ex = expf(qr[k] - mx); qr[k] = ex; sum += ex; }
        const float inv = 1.0f / sum;
        float t = 0.f;
        for (int k = 0; k < K; ++k) { qr[k] *= inv; t += qr[k]; }
        const float rn = 1.0f / (t + 1e-10f);
        for (int k = 0; k < K; ++k) qr[k] *= rn;
      }
    }
    __syncthreads();
    if (it + 1 < nit) {   // the decision pose of the last iteration is never consumed
      float* vo = s.v + it * KD;
      for (int c = tid; c < KD; c += RT_THREADS) {
        const int k = c >> 6;
        float acc = 0.f;
        if (it == 0) {
          for (int r = 0; r < 10; ++r) acc += s.u[r * KD + c];
          acc *= invK;
        } else {
          for (int r = 0; r < 10; ++r) acc = fmaf(q[r * K + k] * s.act[r], s.u[r * KD + c], acc);
        }
        vo[c] = acc;
      }
      __syncthreads();
    }
  }
  // R = q*mask / clamp_min(sum_r q*mask, 1e-10)   (route_given_pheno)
  const float* ql = s.q + (nit - 1) * 10 * K;
  if (tid < K) {
    float den = 0.f;
    for (int r = 0; r < 10; ++r) den += ql[r * K + tid] * s.rm[r];
    den = fmaxf(den, 1e-10f);
    for (int r = 0; r < 10; ++r) s.Rn[r * K + tid] = ql[r * K + tid] * s.rm[r] / den;
  }
  __syncthreads();
  for (int i = tid; i < K * 32; i += RT_THREADS) {
    const int k = i >> 5, p = i & 31;
    float acc = 0.f;
    for (int r = 0; r < 10; ++r) {
      const float cr = (a.d.variant == MMR_VARIANT_PHENO) ? s.alpha[r] : 1.f;
      acc = fmaf(s.Rn[r * K + k] * cr, s.pose[r * 32 + p], acc);
    }
    s.dp[i] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(RT_THREADS) routing_fwd_kernel(RoutingArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int K = a.d.K;
  RtSmem s = rt_carve(smem, K, a.d.num_routing, false);
  rt_build_G(a, s);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = blockIdx.x; b < a.d.B; b += gridDim.x) {
    rt_forward(a, s, b);
    for (int k = warp; k < K; k += RT_THREADS / 32) {
      float acc = s.dp[k * 32 + lane] * s.G[k * 32 + lane];
      acc = warp_sum(acc);
      if (lane == 0) a.logits[(size_t)b * K + k] = acc + a.p.bias[k];
    }
    if (tid < 10) a.alpha[(size_t)b * 10 + tid] = s.alpha[tid];
    if (a.R)
      for (int i = tid; i < 10 * K; i += RT_THREADS) a.R[(size_t)b * 10 * K + i] = s.Rn[i];
  }
}

// 32 per-lane partial sums -> lane i ends with the warp-wide total of index i (31 shuffles).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float send = (lane & 16) ? v[i] : v[i + 16];
    const float keep = (lane & 16) ? v[i + 16] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = (lane & 8) ? v[i] : v[i + 8];
    const float keep = (lane & 8) ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 4) ? v[i] : v[i + 4];
    const float keep = (lane & 4) ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 2) ? v[i] : v[i + 2];
    const float keep = (lane & 2) ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const float send = (lane & 1) ? v[0] : v[1];
    const float keep = (lane & 1) ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

__global__ void __launch_bounds__(RT_THREADS) routing_bwd_kernel(RoutingArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int K = a.d.K, KD = K * 64, nit = a.d.num_routing;
  RtSmem s = rt_carve(smem, K, nit, true);
  rt_build_G(a, s);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool pheno = a.d.variant == MMR_VARIANT_PHENO;
  const bool has_mask = a.route_mask != nullptr;
  const float scale = 0.125f, invK = 1.0f / (float)K;
  // per-CTA accumulators for dG / dbias (flushed once at the end)
  float accG[4] = {0.f, 0.f, 0.f, 0.f};   // element i = tid + 256*j of [K][32]  (K*32 <= 1024)
  float accB = 0.f;                         // tid < K
  for (int b = blockIdx.x; b < a.d.B; b += gridDim.x) {
    rt_forward(a, s, b);
    const float* dl = a.d_logits + (size_t)b * K;
    // (a,b) head: dG, dbias, ddp = dlogit[k]*G[k][p]
    for (int j = 0; j < 4; ++j) {
      const int i = tid + 256 * j;
      if (i < K * 32) {
        const float g = dl[i >> 5];
        accG[j] = fmaf(g, s.dp[i], accG[j]);
        s.ddp[i] = g * s.G[i];
      }
    }
    if (tid < K) accB += dl[tid];
    for (int i = tid; i < 320; i += RT_THREADS) { s.dpose[i] = 0.f; a.poses_m[(size_t)b * 320 + i] = s.pose[i]; }
    __syncthreads();
    // (c) t_rk = sum_p ddp[k][p]*pose[r][p];  dRt = dR + c_r*t_rk  (stored in ds[nit-1] temporarily)
    // The dv slab of the last iteration is never used by the iteration backward (the last
    // decision pose is not consumed), so it serves as scratch for t_rk.
    float* dRt = s.ds + (nit - 1) * 10 * K;
    float* tr = s.dv + (nit - 1) * KD;
    for (int o = tid; o < 10 * K; o += RT_THREADS) {
      const int r = o / K, k = o % K;
      float t = 0.f;
      for (int p = 0; p < 32; ++p) t = fmaf(s.ddp[k * 32 + p], s.pose[r * 32 + p], t);
      const float cr = pheno ? s.alpha[r] : 1.f;
      const float dr = a.d_R ? a.d_R[(size_t)b * 10 * K + o] : 0.f;
      tr[o] = t;
      dRt[o] = dr + cr * t;
    }
    __syncthreads();
    // dpose from the final aggregation, dalpha (pheno)
    for (int i = tid; i < 320; i += RT_THREADS) {
      const int r = i >> 5, p = i & 31;
      const float cr = pheno ? s.alpha[r] : 1.f;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(s.Rn[r * K + k], s.ddp[k * 32 + p], acc);
      s.dpose[i] = cr * acc;
    }
    if (tid < 10) {
      float da = 0.f;
      if (pheno)
        for (int k = 0; k < K; ++k) da = fmaf(s.Rn[tid * K + k], tr[tid * K + k], da);
      s.misc[tid] = da;          // d alpha
      s.misc[16 + tid] = 0.f;    // d act (filled by the iterations)
    }
    __syncthreads();
    // (d) through R = qm / den
    const float* ql = s.q + (nit - 1) * 10 * K;
    if (tid < K) {
      const int k = tid;
      float den = 0.f, dot = 0.f;
      for (int r = 0; r < 10; ++r) { den += ql[r * K + k] * s.rm[r]; dot = fmaf(dRt[r * K + k], s.Rn[r * K + k], dot); }
      const bool clamped = den < 1e-10f;
      const float dd = clamped ? 1e-10f : den;
      for (int r = 0; r < 10; ++r) {
        const float g = clamped ? dRt[r * K + k] : (dRt[r * K + k] - dot);
        dRt[r * K + k] = g / dd * s.rm[r];     // now holds dq of the last iteration
      }
    }
    __syncthreads();
    // (e) agreement iterations, last to first.  ds[it] holds dq_it on entry and ds_it on exit.
    for (int it = nit - 1; it >= 1; --it) {
      float* dq = s.ds + it * 10 * K;
      const float* q = s.q + it * 10 * K;
      if (tid < 10) {
        const int r = tid;
        float aa = 0.f, sq = 0.f;
        for (int k = 0; k < K; ++k) { aa = fmaf(dq[r * K + k], q[r * K + k], aa); sq += q[r * K + k]; }
        const float T = 1.0f;   // sum of the softmax (== 1 up to rounding) + 1e-10
        float bb = 0.f;
        for (int k = 0; k < K; ++k) bb = fmaf((dq[r * K + k] - aa) / T, q[r * K + k], bb);
        for (int k = 0; k < K; ++k) dq[r * K + k] = q[r * K + k] * ((dq[r * K + k] - aa) / T - bb);
        (void)sq;
      }
      __syncthreads();
      // dv_{it-1}[c] = scale * sum_r ds[r][k] * u[r][c]
      float* dvp = s.dv + (it - 1) * KD;
      for (int c = tid; c < KD; c += RT_THREADS) {
        const int k = c >> 6;
        float acc = 0.f;
        for (int r = 0; r < 10; ++r) acc = fmaf(dq[r * K + k], s.u[r * KD + c], acc);
        dvp[c] = acc * scale;
      }
      __syncthreads();
      if (it - 1 >= 1) {
        // v_{it-1} = sum_r q_{it-1}*act*u  ->  dq_{it-1}[r][k] = act[r]*w_rk, dact[r] += sum_k q*w_rk
        float* dqp = s.ds + (it - 1) * 10 * K;
        const float* qp = s.q + (it - 1) * 10 * K;
        for (int o = warp; o < 10 * K; o += RT_THREADS / 32) {
          const int r = o / K, k = o % K;
          const float* uu = s.u + r * KD + k * 64;
          float w = uu[lane] * dvp[k * 64 + lane] + uu[lane + 32] * dvp[k * 64 + lane + 32];
          w = warp_sum(w);
          if (lane == 0) {
            dqp[o] = s.act[r] * w;
            atomicAdd(&s.misc[16 + r], qp[o] * w);
          }
        }
        __syncthreads();
      }
    }
    // (f) du[r][c] and dpose[r][a] += sum_c du[r][c]*w[r][a][c]
    for (int r = 0; r < 10; ++r) {
      float part[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) part[i] = 0.f;
      for (int c = tid; c < KD; c += RT_THREADS) {
        const int k = c >> 6;
        float du = 0.f;
        if (nit >= 2) du = invK * s.dv[c];
        for (int it = 1; it < nit; ++it) {
          du = fmaf(s.ds[it * 10 * K + r * K + k], scale * s.v[(it - 1) * KD + c], du);
          if (it < nit - 1) du = fmaf(s.q[it * 10 * K + r * K + k] * s.act[r], s.dv[it * KD + c], du);
        }
        a.du[((size_t)b * 10 + r) * KD + c] = du;
        const float* w = a.p.caps_w + (size_t)r * 32 * KD + c;
#pragma unroll
        for (int aa = 0; aa < 32; ++aa) part[aa] = fmaf(du, w[(size_t)aa * KD], part[aa]);
      }
      const float tot = warp_transpose_reduce(part, lane);
      atomicAdd(&s.dpose[r * 32 + lane], tot);
    }
    __syncthreads();
    // (g) activation chain and projector data-gradient
    if (tid < 10) {
      const int r = tid;
      const float rm = s.rm[r];
      float dal = s.misc[r] + (pheno ? s.misc[16 + r] : 0.f);   // d alpha (alpha = a3*rm)
      float g = has_mask ? dal * rm : dal;                       // d a3
      if (a.d.from_poses) {
        if (a.d_acts) a.d_acts[(size_t)b * 10 + r] = g;
        s.misc[32 + r] = 0.f;
      } else {
        if (a.d.detach_priors) g = 0.f;
        const bool keep = has_mask ? (rm != 0.f) : true;
        if (keep) { const float x = s.a2[r]; if (x < a.d.prior_floor || x > a.d.prior_ceiling) g = 0.f; }
        if (a.d.act_temperature != 1.0f && has_mask && keep) {
          const float a1 = s.a0[r] * rm;
          if (a1 < 1e-6f || a1 > 1.0f - 1e-6f) g = 0.f;
          else { const float y = s.a2[r]; g *= y * (1.0f - y) / a.d.act_temperature / (a1 * (1.0f - a1)); }
        }
        if (has_mask) g *= rm;
        if (a.acts_override) g = 0.f;
        else g *= s.a0[r] * (1.0f - s.a0[r]);
        s.misc[32 + r] = g;   // d (activation logit)
      }
    }
    __syncthreads();
    if (a.d.from_poses) {
      if (a.d_poses)
        for (int i = tid; i < 320; i += RT_THREADS) a.d_poses[(size_t)b * 320 + i] = s.dpose[i] * s.rm[i >> 5];
    } else {
      for (int i = tid; i < 330; i += RT_THREADS) {
        const int r = i / 33, j = i % 33;
        const float g = (j < 32) ? s.dpose[r * 32 + j] * s.rm[r] : s.misc[32 + r];
        a.dpc[(size_t)b * 330 + i] = g;
        s.e[i] = g;     // route embeddings are no longer needed: reuse as dpc[10][33]
      }
      __syncthreads();
      if (a.d_route_embs) {
        for (int r = 0; r < 10; ++r) {
          const float* w = a.p.proj_w[r] + tid;
          float acc = 0.f;
#pragma unroll 3
          for (int j = 0; j < 33; ++j) acc = fmaf(s.e[r * 33 + j], w[(size_t)j * 256], acc);
          a.d_route_embs[(size_t)r * a.d.emb_route_stride + (size_t)b * a.d.emb_batch_stride + tid] = acc;
        }
      }
    }
    __syncthreads();
  }
  for (int j = 0; j < 4; ++j) {
    const int i = tid + 256 * j;
    if (i < K * 32 && a.dG) atomicAdd(a.dG + i, accG[j]);
  }
  if (tid < K && a.dbias) atomicAdd(a.dbias + tid, accB);
}

// dG[k][p] -> d pose_to_mc[m][p] += sum_k dG[k][p]*emb[k][m];  d emb[k][m] += sum_p dG[k][p]*Wmc[m][p]
__global__ void routing_head_grads_kernel(const float* dG, const float* pose_to_mc, const float* embedding, int K,
                                          float* d_pose_to_mc, float* d_embedding) {
  const int tid = threadIdx.x;
  if (d_pose_to_mc)
    for (int i = tid; i < MC * PC; i += blockDim.x) {
      const int m = i / PC, p = i % PC;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc = fmaf(dG[k * 32 + p], embedding[k * MC + m], acc);
      d_pose_to_mc[i] += acc;
    }
  if (d_embedding)
    for (int i = tid; i < K * MC; i += blockDim.x) {
      const int k = i / MC, m = i % MC;
      float acc = 0.f;
      for (int p = 0; p < PC; ++p) acc = fmaf(dG[k * 32 + p], pose_to_mc[m * PC + p], acc);
      d_embedding[i] += acc;
    }
}

}  // namespace mmr
