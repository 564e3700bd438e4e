// Capsule routing, reduced-precision mode, round-2 formulation: the dense contractions and the agreement iterations
// are separate launches so that each gets the parallelism it needs.
//
//   rs_project_kernel   projector  [16 patients x 256] . [256 x 33] per route on mma.sync (M = 16 is a FULL tile; the
//                       tile-of-4 kernel in routing.cuh used 4 of 16 rows and re-streamed the weights per 4 patients)
//   rs_votes_kernel     votes u[b][r][c] = sum_a pose_m[b][r][a] w[r][a][c] as a GEMM over (16-patient tile, column slice),
//                       written once as saturating fp16 to a global scratch that stays L2-resident (16 MB at B = 512, K = 25)
//   rs_iterate_*_kernel the agreement iterations of ONE patient on 1 / 2 / 4 warps: lane = label, warps (and, for small K, lane
//                       groups) split the 64 vote dimensions.  Everything that is per label -- agreement dots, decision
//                       poses v_it, their gradients -- stays in registers; softmax over labels is a shuffle reduction; the
//                       only exchange between the warps of a patient is the 10-value partial dot, through a ping-pong
//                       shared-memory slot and a NAMED barrier (bar.sync id, 32 NW), never a CTA-wide barrier.  The votes
//                       of the patient are staged once in shared memory (cp.async, 36 KB at K = 25), so 4 patients are
//                       resident per SM and B = 512 is a single wave.
//   rs_dpose_kernel     d pose = du . w^T (mma.sync, M = 16 patients, split over the reduction) + the final-aggregation
//                       term, then the projector data gradient for the same (tile, route).
//
// Semantics are those of routing.cuh (same reference lines): RoutePrimaryProjector.forward routing_and_heads.py:111-121,
// forward_capsule_from_route_dict :314-352, CapsuleMortalityHead.forward Mort :194-268 / Pheno :194-272, CapsuleFC.forward
// capsule_layers.py:75-117.  Used when vote_dtype = MMR_DTYPE_BF16, the fp16 weight copies exist and num_routing <= 3;
// everything else (fp32 parity mode, num_routing = 4) stays on routing.cuh.
#pragma once
#include "routing.cuh"

namespace mmr {

#ifndef RS_BWD_MINB
#define RS_BWD_MINB 4
#endif
constexpr int RS_NIT = 3;           // largest num_routing of this path (register-resident v_it / dv_it)
constexpr int RS_UP = 72;           // pitch of a staged vote row in halves (64 + 8: conflict-free 16-byte row reads per label)
constexpr int RS_GCOPIES = 16;      // the head-gradient atomics of a CTA go to copy blockIdx.x % 16 (444-way -> 28-way contention)
constexpr float RS_DU_SCALE = 1024.f;   // du is scaled into the fp16 normal range before the d pose contraction

struct RsScratch {
  float* pose;      // [B,10,32] projector poses (unmasked)
  float* zl;        // [B,10]    activation logits
  float* G;         // [K,32]    embedding @ pose_to_mc
  __half* votes;    // [B,10,K*64]
  float* qs;        // [B,RS_NIT-1,10,32] routing coefficients q_1 .. q_{nit-1} of the forward call (read by the backward; may be null)
  float* dposeA;    // bwd [B,10,32]: final-aggregation part of d pose_m
  float* dGc;       // bwd [RS_GCOPIES][(K+1)*32]: per-copy accumulators of dG (rows < K) and d bias (row K), zeroed by the host
};

__device__ __forceinline__ void rs_mma16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// A fragments (two k-steps) of 8 consecutive fp32 values through the read-only path, optionally scaled
__device__ __forceinline__ void rs_afrag_ldg(const float* x8, bool valid, float sc, uint32_t (&A)[4]) {
  A[0] = A[1] = A[2] = A[3] = 0u;
  if (valid) {
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(x8)), x1 = __ldg(reinterpret_cast<const float4*>(x8 + 4));
    A[0] = rt_pack2(x0.x * sc, x0.y * sc); A[1] = rt_pack2(x0.z * sc, x0.w * sc);
    A[2] = rt_pack2(x1.x * sc, x1.y * sc); A[3] = rt_pack2(x1.z * sc, x1.w * sc);
  }
}
__device__ __forceinline__ void rs_cp16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void rs_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void rs_group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- projector: grid (ceil(B / 16), 11), 128 threads; blockIdx.y = route, the extra row builds G ---------------------
__global__ void __launch_bounds__(128) rs_project_kernel(RoutingArgs a, RsScratch s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  if (blockIdx.y == 10) {
    if (blockIdx.x != 0) return;
    for (int i = tid; i < a.d.K * 32; i += 128) {
      const int k = i >> 5, p = i & 31;
      float acc = 0.f;
      for (int m = 0; m < MC; ++m) acc = fmaf(a.p.pose_to_mc[m * PC + p], a.p.embedding[k * MC + m], acc);
      s.G[i] = acc;
    }
    return;
  }
  if (a.d.from_poses) return;
  __shared__ float part[4][16][41];
  const int r = blockIdx.y, b0 = blockIdx.x * 16, B = a.d.B;
  const bool vl = b0 + g < B, vh = b0 + g + 8 < B;
  const float* el = a.route_embs + (size_t)r * a.d.emb_route_stride + (size_t)(vl ? b0 + g : 0) * a.d.emb_batch_stride + 8 * t;
  const float* eh = a.route_embs + (size_t)r * a.d.emb_route_stride + (size_t)(vh ? b0 + g + 8 : 0) * a.d.emb_batch_stride + 8 * t;
  const uint4* pw = reinterpret_cast<const uint4*>(a.p.proj_w_f16);
  float acc[5][4];
#pragma unroll
  for (int n = 0; n < 5; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {       // each warp owns two of the eight 32-wide slices of the reduction
    const int kk = warp * 2 + j;
    uint32_t Al[4], Ah[4];
    rs_afrag_ldg(el + 32 * kk, vl, 1.f, Al);
    rs_afrag_ldg(eh + 32 * kk, vh, 1.f, Ah);
#pragma unroll
    for (int n = 0; n < 5; ++n) {
      const uint4 q = __ldg(pw + ((size_t)(r * 40 + n * 8 + g) * 32 + kk * 4 + t));
      rs_mma16(acc[n], Al[0], Ah[0], Al[1], Ah[1], q.x, q.y);
      rs_mma16(acc[n], Al[2], Ah[2], Al[3], Ah[3], q.z, q.w);
    }
  }
#pragma unroll
  for (int n = 0; n < 5; ++n) {
    part[warp][g][n * 8 + 2 * t] = acc[n][0]; part[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
    part[warp][g + 8][n * 8 + 2 * t] = acc[n][2]; part[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
  }
  __syncthreads();
  for (int i = tid; i < 16 * 33; i += 128) {
    const int p = i / 33, j = i % 33, b = b0 + p;
    if (b >= B) continue;
    const float v = part[0][p][j] + part[1][p][j] + part[2][p][j] + part[3][p][j] + __ldg(a.p.proj_b[r] + j);
    if (j < 32) {
      s.pose[((size_t)b * 10 + r) * 32 + j] = v;
      if (a.poses_out) a.poses_out[((size_t)b * 10 + r) * 32 + j] = v;
    } else {
      s.zl[(size_t)b * 10 + r] = v;
      if (a.acts_out) a.acts_out[(size_t)b * 10 + r] = 1.0f / (1.0f + expf(-v));
    }
  }
}

// ---- votes: grid (ceil(B / 16), S), 256 threads; a warp owns 16-column units and walks the 10 routes -----------------
// The two n-tiles of a unit take their B columns in the order 4 (g >> 1) + 2 j + (g & 1), so that a lane ends up with four
// CONSECUTIVE columns of its rows (one 8-byte store per row instead of two 4-byte ones).
__device__ __forceinline__ uint2 rs_pack4(float x0, float x1, float x2, float x3) {
  uint2 o;
  o.x = rt_pack2(x0, x1); o.y = rt_pack2(x2, x3);
  return o;
}
__global__ void __launch_bounds__(256) rs_votes_kernel(RoutingArgs a, RsScratch s) {
  __shared__ __align__(16) float pm[16][324];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int B = a.d.B, KD = a.d.K * 64, b0 = blockIdx.x * 16;
  const float* psrc = a.d.from_poses ? a.poses_in : s.pose;
  {   // 16 x 320 masked poses: all loads of a thread are issued before the first store (read-only path)
    float4 v[5]; float rm[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = tid + 256 * i, p = idx / 80, j4 = idx % 80, b = b0 + p;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f); rm[i] = 1.f;
      if (b < B) {
        v[i] = __ldg(reinterpret_cast<const float4*>(psrc + (size_t)b * 320) + j4);
        if (a.route_mask) rm[i] = __ldg(a.route_mask + (size_t)b * 10 + (j4 >> 3));
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int idx = tid + 256 * i, p = idx / 80, j4 = idx % 80;
      const float4 o = make_float4(v[i].x * rm[i], v[i].y * rm[i], v[i].z * rm[i], v[i].w * rm[i]);
      *reinterpret_cast<float4*>(&pm[p][4 * j4]) = o;
    }
  }
  __syncthreads();
  const int NU = KD / 16;
  const int u_lo = (int)((long long)NU * blockIdx.y / gridDim.y), u_hi = (int)((long long)NU * (blockIdx.y + 1) / gridDim.y);
  const uint4* wt = reinterpret_cast<const uint4*>(a.p.caps_wt_f16);
  const bool vl = b0 + g < B, vh = b0 + g + 8 < B;
  for (int un = u_lo + warp; un < u_hi; un += 8) {
    uint4 q[10][2];
#pragma unroll
    for (int r = 0; r < 10; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = 16 * un + 4 * (g >> 1) + 2 * j + (g & 1);
        q[r][j] = __ldg(wt + ((size_t)r * KD + c) * 4 + t);
      }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t Al[4], Ah[4];
      rt_afrag(&pm[g][r * 32 + 8 * t], true, Al);
      rt_afrag(&pm[g + 8][r * 32 + 8 * t], true, Ah);
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      rs_mma16(c0, Al[0], Ah[0], Al[1], Ah[1], q[r][0].x, q[r][0].y);
      rs_mma16(c0, Al[2], Ah[2], Al[3], Ah[3], q[r][0].z, q[r][0].w);
      rs_mma16(c1, Al[0], Ah[0], Al[1], Ah[1], q[r][1].x, q[r][1].y);
      rs_mma16(c1, Al[2], Ah[2], Al[3], Ah[3], q[r][1].z, q[r][1].w);
      if (vl) *reinterpret_cast<uint2*>(s.votes + ((size_t)(b0 + g) * 10 + r) * KD + 16 * un + 4 * t) = rs_pack4(c0[0], c0[1], c1[0], c1[1]);
      if (vh) *reinterpret_cast<uint2*>(s.votes + ((size_t)(b0 + g + 8) * 10 + r) * KD + 16 * un + 4 * t) = rs_pack4(c0[2], c0[3], c1[2], c1[3]);
    }
  }
}

// ---- agreement iterations ---------------------------------------------------------------------------------------------
template <int KP> struct RsCfg {
  static constexpr int S = 32 / KP;                                   // vote-dimension segments inside a warp
  static constexpr int NW = KP == 32 ? 4 : (KP == 16 ? 2 : 1);        // warps per patient
  static constexpr int NSEG = S * NW;
  static constexpr int DPL = 64 / NSEG;                               // vote dimensions per lane: 16, 16, 16, 8, 4
  static constexpr int PPL = 32 / NSEG;                               // pose columns per lane:     8,  8,  8, 4, 2
  static constexpr int PPC = 4 / NW;                                  // patients per 128-thread CTA
};
// per-patient table (floats): a0 a2 a3 alpha act rm c dal dact  (16 each)
constexpr int RS_TAB = 9 * 16;
__host__ __device__ inline size_t rs_patient_smem(int K, int KP, int NW) {
  return (size_t)10 * K * RS_UP * 2 + (size_t)2 * NW * 10 * KP * 4 + 320 * 4 + RS_TAB * 4;
}

template <int N> __device__ __forceinline__ void rs_ld_u(const __half* p, float (&f)[N]);
template <> __device__ __forceinline__ void rs_ld_u<16>(const __half* p, float (&f)[16]) {
  const uint4 x0 = *reinterpret_cast<const uint4*>(p), x1 = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
template <> __device__ __forceinline__ void rs_ld_u<8>(const __half* p, float (&f)[8]) {
  const uint4 x0 = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
template <> __device__ __forceinline__ void rs_ld_u<4>(const __half* p, float (&f)[4]) {
  const uint2 x0 = *reinterpret_cast<const uint2*>(p);
  const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&x0.x));
  const float2 v1 = __half22float2(*reinterpret_cast<const __half2*>(&x0.y));
  f[0] = v0.x; f[1] = v0.y; f[2] = v1.x; f[3] = v1.y;
}

template <int KP> __device__ __forceinline__ float rs_ksum(float v) {     // over the labels (lane = seg * KP + k)
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int KP> __device__ __forceinline__ float rs_kmax(float v) {
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// state of one patient group (NW warps) inside the iterate kernels
template <int KP> struct RsGroup {
  using C = RsCfg<KP>;
  __half* u; float* xp; float* pm; float* tab;
  int buf, wp, k, sg, seg, lane, bar; bool kv;
  __device__ __forceinline__ void sync() {
    if (C::NW > 1) rs_group_bar(bar, C::NW * 32); else __syncwarp();
  }
  // sum over the vote-dimension segments of an array every lane holds for its label
  template <int N> __device__ __forceinline__ void seg_reduce(float (&x)[N]) {
#pragma unroll
    for (int o = KP; o < 32; o <<= 1)
#pragma unroll
      for (int i = 0; i < N; ++i) x[i] += __shfl_xor_sync(0xffffffffu, x[i], o);
    if (C::NW > 1) {
      float* mine = xp + (size_t)((buf * C::NW + wp) * 10) * KP;
      if (sg == 0)
#pragma unroll
        for (int i = 0; i < N; ++i) mine[i * KP + k] = x[i];
      rs_group_bar(bar, C::NW * 32);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < C::NW; ++w) acc += xp[(size_t)((buf * C::NW + w) * 10 + i) * KP + k];
        x[i] = acc;
      }
      buf ^= 1;
    }
  }
  __device__ __forceinline__ const __half* urow(int r, int K) const { return u + (size_t)(r * K + (kv ? k : K - 1)) * RS_UP + seg * C::DPL; }
};

// stage the patient's votes, masked poses and activation chain (routing_and_heads.py:314-352)
template <int KP>
__device__ __forceinline__ void rs_stage(const RoutingArgs& a, const RsScratch& s, RsGroup<KP>& G, int b) {
  using C = RsCfg<KP>;
  const int K = a.d.K, KD = K * 64, gt = G.wp * 32 + G.lane;
  const __half* src = s.votes + (size_t)b * 10 * KD;
  for (int c = gt; c < 10 * K * 8; c += C::NW * 32) rs_cp16(G.u + (size_t)(c >> 3) * RS_UP + (c & 7) * 8, src + (size_t)c * 8);
  const bool has_mask = a.route_mask != nullptr;
  const float* psrc = a.d.from_poses ? a.poses_in : s.pose;
  for (int i = gt; i < 320; i += C::NW * 32) {
    const float rm = has_mask ? __ldg(a.route_mask + (size_t)b * 10 + (i >> 5)) : 1.f;
    G.pm[i] = __ldg(psrc + (size_t)b * 320 + i) * rm;
  }
  if (G.wp == 0 && G.lane < 10) {
    const int r = G.lane;
    const float rm = has_mask ? a.route_mask[(size_t)b * 10 + r] : 1.f;
    float a0, a2, a3, alpha;
    if (!a.d.from_poses) {
      a0 = 1.0f / (1.0f + expf(-s.zl[(size_t)b * 10 + r]));
      if (a.acts_override) a0 = a.acts_override[(size_t)b * 10 + r];
      float x = has_mask ? a0 * rm : a0;
      const bool keep = has_mask ? (rm != 0.f) : true;
      if (a.d.act_temperature != 1.0f && has_mask && keep) {
        const float xc = fminf(fmaxf(x, 1e-6f), 1.0f - 1e-6f);
        const float lg = (logf(xc) - log1pf(-xc)) / a.d.act_temperature;
        x = 1.0f / (1.0f + expf(-lg));
      }
      a2 = x;
      if (keep) x = fminf(fmaxf(x, a.d.prior_floor), a.d.prior_ceiling);
      a3 = x;
      alpha = has_mask ? x * rm : x;
    } else {
      const float x = a.acts_in[(size_t)b * 10 + r];
      a0 = a2 = a3 = x;
      alpha = has_mask ? x * rm : x;
    }
    const bool pheno = a.d.variant == MMR_VARIANT_PHENO;
    G.tab[r] = a0; G.tab[16 + r] = a2; G.tab[32 + r] = a3; G.tab[48 + r] = alpha;
    G.tab[64 + r] = pheno ? alpha : rm;     // routing activation
    G.tab[80 + r] = rm;
    G.tab[96 + r] = pheno ? alpha : 1.f;    // coefficient of the final aggregation
  }
  rs_cp_wait();
  G.sync();
}

// forward iterations; on exit q = last routing coefficients; V (when STORE) keeps v_0 .. v_{nit-2}, Q keeps q_1 .. q_{nit-1}
// qsave (forward kernel): q_it is written there; qload (backward after a forward call with scratch): q_it is read back instead
// of recomputing the agreement dots and the softmax.  Both point at this patient's [RS_NIT-1][10][32] block or are null.
template <int KP, bool STORE>
__device__ __forceinline__ void rs_forward(const RoutingArgs& a, RsGroup<KP>& G, float (&q)[10],
                                           float (&V)[RS_NIT - 1][RsCfg<KP>::DPL], float (&Q)[RS_NIT - 1][10],
                                           float* qsave, const float* qload) {
  using C = RsCfg<KP>;
  constexpr int DPL = C::DPL;
  const int K = a.d.K, nit = a.d.num_routing;
  const float invK = 1.0f / (float)K, scale = 0.125f;
#pragma unroll
  for (int r = 0; r < 10; ++r) q[r] = G.kv ? invK : 0.f;
  float v[DPL];
#pragma unroll
  for (int d = 0; d < DPL; ++d) v[d] = 0.f;
  if (nit > 1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      float f[DPL];
      rs_ld_u<DPL>(G.urow(r, K), f);
#pragma unroll
      for (int d = 0; d < DPL; ++d) v[d] += f[d];
    }
#pragma unroll
    for (int d = 0; d < DPL; ++d) v[d] *= invK;
  }
#pragma unroll
  for (int it = 1; it < RS_NIT; ++it) {
    if (it < nit) {
      if (STORE) {
#pragma unroll
        for (int d = 0; d < DPL; ++d) V[it - 1][d] = v[d];
      }
      if (qload) {
#pragma unroll
        for (int r = 0; r < 10; ++r) q[r] = G.kv ? qload[((it - 1) * 10 + r) * 32 + G.k] : 0.f;
      } else {
        float x[10];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          float f[DPL];
          rs_ld_u<DPL>(G.urow(r, K), f);
          float acc = 0.f;
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc = fmaf(f[d], v[d], acc);
          x[r] = acc;
        }
        G.template seg_reduce<10>(x);
#pragma unroll
        for (int r = 0; r < 10; ++r) {       // softmax over the labels, renormalised (capsule_layers.py:96-100)
          const float xv = G.kv ? scale * x[r] : -INFINITY;
          const float mx = rs_kmax<KP>(xv);
          const float ex = G.kv ? expf(xv - mx) : 0.f;
          const float pr = ex * (1.0f / rs_ksum<KP>(ex));
          const float tt = rs_ksum<KP>(pr);
          q[r] = pr * (1.0f / (tt + 1e-10f));
        }
        if (qsave && G.kv && G.seg == 0)
#pragma unroll
          for (int r = 0; r < 10; ++r) qsave[((it - 1) * 10 + r) * 32 + G.k] = q[r];
      }
      if (STORE) {
#pragma unroll
        for (int r = 0; r < 10; ++r) Q[it - 1][r] = q[r];
      }
      if (it + 1 < nit) {
#pragma unroll
        for (int d = 0; d < DPL; ++d) v[d] = 0.f;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          float f[DPL];
          rs_ld_u<DPL>(G.urow(r, K), f);
          const float cf = q[r] * G.tab[64 + r];
#pragma unroll
          for (int d = 0; d < DPL; ++d) v[d] = fmaf(cf, f[d], v[d]);
        }
      }
    }
  }
}

template <int KP> __device__ __forceinline__ void rs_group_init(RsGroup<KP>& G, uint8_t* smem, int K) {
  using C = RsCfg<KP>;
  const int warp = threadIdx.x >> 5, pg = warp / C::NW;
  uint8_t* base = smem + (size_t)pg * ((rs_patient_smem(K, KP, C::NW) + 15) / 16 * 16);
  G.u = reinterpret_cast<__half*>(base); base += (size_t)10 * K * RS_UP * 2;
  G.xp = reinterpret_cast<float*>(base); base += (size_t)2 * C::NW * 10 * KP * 4;
  G.pm = reinterpret_cast<float*>(base); base += 320 * 4;
  G.tab = reinterpret_cast<float*>(base);
  G.buf = 0; G.wp = warp % C::NW; G.lane = threadIdx.x & 31;
  G.k = G.lane % KP; G.sg = G.lane / KP; G.seg = G.wp * C::S + G.sg;
  G.bar = 1 + pg; G.kv = G.k < K;
}

template <int KP>
__global__ void __launch_bounds__(128, 4) rs_iterate_fwd_kernel(RoutingArgs a, RsScratch s) {
  using C = RsCfg<KP>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = a.d.K, B = a.d.B;
  RsGroup<KP> G;
  rs_group_init<KP>(G, smem_raw, K);
  const int pg = (threadIdx.x >> 5) / C::NW;
  for (int b = blockIdx.x * C::PPC + pg; b < B; b += gridDim.x * C::PPC) {
    rs_stage<KP>(a, s, G, b);
    float q[10], V[RS_NIT - 1][C::DPL], Q[RS_NIT - 1][10];
    rs_forward<KP, false>(a, G, q, V, Q, s.qs ? s.qs + (size_t)b * (RS_NIT - 1) * 320 : nullptr, nullptr);
    // R = q mask / clamp_min(sum_r q mask, 1e-10)   (route_given_pheno)
    float den = 0.f;
#pragma unroll
    for (int r = 0; r < 10; ++r) den = fmaf(q[r], G.tab[80 + r], den);
    den = fmaxf(den, 1e-10f);
    float Rn[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) Rn[r] = q[r] * G.tab[80 + r] / den;
    if (a.R && G.kv && G.seg == 0)
#pragma unroll
      for (int r = 0; r < 10; ++r) a.R[(size_t)b * 10 * K + r * K + G.k] = Rn[r];
    float lg[1] = {0.f};
#pragma unroll
    for (int j = 0; j < C::PPL; ++j) {
      const int p = G.seg * C::PPL + j;
      float dp = 0.f;
#pragma unroll
      for (int r = 0; r < 10; ++r) dp = fmaf(Rn[r] * G.tab[96 + r], G.pm[r * 32 + p], dp);
      lg[0] = fmaf(dp, G.kv ? s.G[G.k * 32 + p] : 0.f, lg[0]);
    }
    G.template seg_reduce<1>(lg);
    if (G.kv && G.seg == 0) a.logits[(size_t)b * K + G.k] = lg[0] + a.p.bias[G.k];
    if (G.wp == 0 && G.lane < 10) a.alpha[(size_t)b * 10 + G.lane] = G.tab[48 + G.lane];
    G.sync();     // the next patient's staging overwrites u / pm / tab
  }
}

template <int KP>
__global__ void __launch_bounds__(128, RS_BWD_MINB) rs_iterate_bwd_kernel(RoutingArgs a, RsScratch s) {
  using C = RsCfg<KP>;
  constexpr int DPL = C::DPL, PPL = C::PPL;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = a.d.K, KD = K * 64, B = a.d.B, nit = a.d.num_routing;
  const bool pheno = a.d.variant == MMR_VARIANT_PHENO, has_mask = a.route_mask != nullptr;
  const float invK = 1.0f / (float)K, scale = 0.125f;
  RsGroup<KP> G;
  rs_group_init<KP>(G, smem_raw, K);
  const int pg = (threadIdx.x >> 5) / C::NW;
  float accG[PPL], accB = 0.f;
#pragma unroll
  for (int j = 0; j < PPL; ++j) accG[j] = 0.f;
  for (int b = blockIdx.x * C::PPC + pg; b < B; b += gridDim.x * C::PPC) {
    rs_stage<KP>(a, s, G, b);
    if (a.poses_m)       // masked poses: operand of the vote-weight gradient
      for (int i = G.wp * 32 + G.lane; i < 320; i += C::NW * 32) a.poses_m[(size_t)b * 320 + i] = G.pm[i];
    float q[10], V[RS_NIT - 1][DPL], Q[RS_NIT - 1][10];
    rs_forward<KP, true>(a, G, q, V, Q, nullptr, s.qs ? s.qs + (size_t)b * (RS_NIT - 1) * 320 : nullptr);
    float den = 0.f;
#pragma unroll
    for (int r = 0; r < 10; ++r) den = fmaf(q[r], G.tab[80 + r], den);
    const bool clamped = den < 1e-10f;
    const float dd = clamped ? 1e-10f : den;
    float Rn[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) Rn[r] = q[r] * G.tab[80 + r] / dd;
    // head: dG += dlogit dp, ddp = dlogit G   (this lane's pose columns)
    const float dl = G.kv ? a.d_logits[(size_t)b * K + G.k] : 0.f;
    if (G.seg == 0) accB += dl;
    float ddp[PPL];
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      const int p = G.seg * PPL + j;
      float dp = 0.f;
#pragma unroll
      for (int r = 0; r < 10; ++r) dp = fmaf(Rn[r] * G.tab[96 + r], G.pm[r * 32 + p], dp);
      accG[j] = fmaf(dl, dp, accG[j]);
      ddp[j] = G.kv ? dl * s.G[G.k * 32 + p] : 0.f;
    }
    // t_rk = sum_p ddp[k][p] pose_m[r][p];  dRt = dR + c_r t_rk
    float t[10];
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < PPL; ++j) acc = fmaf(ddp[j], G.pm[r * 32 + G.seg * PPL + j], acc);
      t[r] = acc;
    }
    G.template seg_reduce<10>(t);
    // final aggregation: d pose_m[r][p] = c_r sum_k Rn[r][k] ddp[k][p]  -> scratch (rs_dpose_kernel adds du . w^T)
    if constexpr (KP == 32) {
      // 80 sums over the 32 labels as a reduce-scatter: every butterfly step halves the values a lane carries
      // (80 + 5 shuffles instead of 400); lane l ends with the sums v = 16 m + (l & 15), v = r * 8 + j
      float x[80];
#pragma unroll
      for (int r = 0; r < 10; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) x[r * 8 + j] = Rn[r] * ddp[j];
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const int o = 1 << st;
        const bool up = (G.lane & o) != 0;
#pragma unroll
        for (int i = 0; i < (80 >> (st + 1)); ++i) {
          // values whose index has bit `st` (of the low four bits) clear / set: x[.. 2i ..], x[.. 2i+1 ..] after compaction
          const float keep = up ? x[2 * i + 1] : x[2 * i], give = up ? x[2 * i] : x[2 * i + 1];
          x[i] = keep + __shfl_xor_sync(0xffffffffu, give, o);
        }
      }
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const float v = x[m] + __shfl_xor_sync(0xffffffffu, x[m], 16);
        // after the four compactions x[m] of lane l is the sum of original index 16 m + (l & 15)
        const int idx = 16 * m + (G.lane & 15), r = idx >> 3, j = idx & 7;
        if (G.lane < 16) s.dposeA[(size_t)b * 320 + r * 32 + G.seg * PPL + j] = v * G.tab[96 + r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < 10; ++r)
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
          const float v = rs_ksum<KP>(Rn[r] * ddp[j]) * G.tab[96 + r];
          if (G.k == 0) s.dposeA[(size_t)b * 320 + r * 32 + G.seg * PPL + j] = v;
        }
    }
    float dal[10], dact[10], dq[10];
    float dot = 0.f;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      dal[r] = pheno ? rs_ksum<KP>(Rn[r] * t[r]) : 0.f;
      dact[r] = 0.f;
      const float dr = (a.d_R && G.kv) ? a.d_R[(size_t)b * 10 * K + r * K + G.k] : 0.f;
      dq[r] = dr + G.tab[96 + r] * t[r];        // dRt
      dot = fmaf(dq[r], Rn[r], dot);
    }
#pragma unroll
    for (int r = 0; r < 10; ++r) dq[r] = (clamped ? dq[r] : dq[r] - dot) / dd * G.tab[80 + r];
    // agreement iterations, last to first
    float DV[RS_NIT - 1][DPL], DS[RS_NIT - 1][10];
#pragma unroll
    for (int i = 0; i < RS_NIT - 1; ++i) {
#pragma unroll
      for (int d = 0; d < DPL; ++d) DV[i][d] = 0.f;
#pragma unroll
      for (int r = 0; r < 10; ++r) DS[i][r] = 0.f;
    }
#pragma unroll
    for (int it = RS_NIT - 1; it >= 1; --it) {
      if (it < nit) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          const float qq = Q[it - 1][r], g = dq[r];
          const float aa = rs_ksum<KP>(g * qq);
          const float bb = rs_ksum<KP>((g - aa) * qq);
          DS[it - 1][r] = qq * ((g - aa) - bb);
        }
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          float f[DPL];
          rs_ld_u<DPL>(G.urow(r, K), f);
#pragma unroll
          for (int d = 0; d < DPL; ++d) DV[it - 1][d] = fmaf(DS[it - 1][r], f[d], DV[it - 1][d]);
        }
#pragma unroll
        for (int d = 0; d < DPL; ++d) DV[it - 1][d] *= scale;
        if (it >= 2) {
          float w[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            float f[DPL];
            rs_ld_u<DPL>(G.urow(r, K), f);
            float acc = 0.f;
#pragma unroll
            for (int d = 0; d < DPL; ++d) acc = fmaf(f[d], DV[it - 1][d], acc);
            w[r] = acc;
          }
          G.template seg_reduce<10>(w);
#pragma unroll
          for (int r = 0; r < 10; ++r) {
            dq[r] = G.tab[64 + r] * w[r];
            dact[r] += rs_ksum<KP>(G.kv ? Q[it - 2][r] * w[r] : 0.f);
          }
        }
      }
    }
    // du[r][k][d] (fp32, operand of the vote-weight gradient and of rs_dpose_kernel)
    if (G.kv) {
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        float f[DPL];
#pragma unroll
        for (int d = 0; d < DPL; ++d) f[d] = nit >= 2 ? invK * DV[0][d] : 0.f;
#pragma unroll
        for (int it = 1; it < RS_NIT; ++it) {
          if (it < nit) {
            const float c1 = DS[it - 1][r] * scale;
#pragma unroll
            for (int d = 0; d < DPL; ++d) f[d] = fmaf(c1, V[it - 1][d], f[d]);
            if (it + 1 < RS_NIT && it < nit - 1) {
              const float c2 = Q[it - 1][r] * G.tab[64 + r];
#pragma unroll
              for (int d = 0; d < DPL; ++d) f[d] = fmaf(c2, DV[it + 1 < RS_NIT ? it : 0][d], f[d]);
            }
          }
        }
        float* dst = a.du + ((size_t)b * 10 + r) * KD + G.k * 64 + G.seg * DPL;
#pragma unroll
        for (int d = 0; d < DPL; d += 4) *reinterpret_cast<float4*>(dst + d) = make_float4(f[d], f[d + 1], f[d + 2], f[d + 3]);
      }
    }
    // activation chain backward (lane = route)
    if (G.wp == 0) {
      if (G.lane == 0) {
#pragma unroll
        for (int r = 0; r < 10; ++r) { G.tab[112 + r] = dal[r]; G.tab[128 + r] = dact[r]; }
      }
      __syncwarp();
      if (G.lane < 10) {
        const int r = G.lane;
        const float rm = G.tab[80 + r];
        const float dalr = G.tab[112 + r] + (pheno ? G.tab[128 + r] : 0.f);
        float g = has_mask ? dalr * rm : dalr;
        if (a.d.from_poses) {
          if (a.d_acts) a.d_acts[(size_t)b * 10 + r] = g;
        } else {
          if (a.d.detach_priors) g = 0.f;
          const bool keep = has_mask ? (rm != 0.f) : true;
          const float a0 = G.tab[r], a2 = G.tab[16 + r];
          if (keep) { if (a2 < a.d.prior_floor || a2 > a.d.prior_ceiling) g = 0.f; }
          if (a.d.act_temperature != 1.0f && has_mask && keep) {
            const float a1 = a0 * rm;
            if (a1 < 1e-6f || a1 > 1.0f - 1e-6f) g = 0.f;
            else g *= a2 * (1.0f - a2) / a.d.act_temperature / (a1 * (1.0f - a1));
          }
          if (has_mask) g *= rm;
          if (a.acts_override) {
            if (a.d_acts) a.d_acts[(size_t)b * 10 + r] = g;
            g = 0.f;
          } else {
            g *= a0 * (1.0f - a0);
          }
          a.dpc[(size_t)b * 330 + r * 33 + 32] = g;
        }
      }
    }
    G.sync();
  }
  float* gc = s.dGc + (size_t)(blockIdx.x % RS_GCOPIES) * (K + 1) * 32;
  if (G.kv) {
#pragma unroll
    for (int j = 0; j < PPL; ++j) atomicAdd(gc + G.k * 32 + G.seg * PPL + j, accG[j]);
    if (G.seg == 0) atomicAdd(gc + K * 32 + G.k, accB);
  }
}

// dG = sum of the copies; d bias += row K of the copies
__global__ void rs_fold_copies_kernel(const float* dGc, int K, float* dG, float* dbias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (K + 1) * 32) return;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < RS_GCOPIES; ++c) acc += dGc[(size_t)c * (K + 1) * 32 + i];
  if (i < K * 32) dG[i] = acc;
  else if (i - K * 32 < K && dbias) dbias[i - K * 32] += acc;
}

// ---- d pose = du . w^T + final aggregation; projector data / bias gradient: grid (ceil(B / 16), 10), 256 threads --------
__global__ void __launch_bounds__(256) rs_dpose_kernel(RoutingArgs a, RsScratch s) {
  __shared__ float part[8][16][33];
  __shared__ float dps[16][44];     // columns 33..39 stay zero (k padding of the tf32 tile); pitch 44: conflict-free A fragments
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int B = a.d.B, KD = a.d.K * 64, r = blockIdx.y, b0 = blockIdx.x * 16;
  const bool vl = b0 + g < B, vh = b0 + g + 8 < B;
  const int G32 = KD / 32;
  const int k_lo = G32 * warp / 8, k_hi = G32 * (warp + 1) / 8;     // the eight warps split the reduction over the votes
  const float* dl = a.du + ((size_t)(vl ? b0 + g : 0) * 10 + r) * KD + 8 * t;
  const float* dh = a.du + ((size_t)(vh ? b0 + g + 8 : 0) * 10 + r) * KD + 8 * t;
  const uint4* w16 = reinterpret_cast<const uint4*>(a.p.caps_w_f16);
  float acc[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll 4
  for (int kk = k_lo; kk < k_hi; ++kk) {
    uint32_t Al[4], Ah[4];
    rs_afrag_ldg(dl + 32 * kk, vl, RS_DU_SCALE, Al);
    rs_afrag_ldg(dh + 32 * kk, vh, RS_DU_SCALE, Ah);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const uint4 q = __ldg(w16 + ((size_t)(r * 32 + n * 8 + g) * KD + 32 * kk + 8 * t) / 8);
      rs_mma16(acc[n], Al[0], Ah[0], Al[1], Ah[1], q.x, q.y);
      rs_mma16(acc[n], Al[2], Ah[2], Al[3], Ah[3], q.z, q.w);
    }
  }
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    part[warp][g][n * 8 + 2 * t] = acc[n][0]; part[warp][g][n * 8 + 2 * t + 1] = acc[n][1];
    part[warp][g + 8][n * 8 + 2 * t] = acc[n][2]; part[warp][g + 8][n * 8 + 2 * t + 1] = acc[n][3];
  }
  if (tid < 16 * 7) dps[tid / 7][33 + tid % 7] = 0.f;
  // operands of the epilogue that do not depend on the contraction: issued before the barrier
  float ex[3], rmv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + 256 * i, p = idx / 33, j = idx % 33, b = b0 + p;
    ex[i] = 0.f; rmv[i] = 1.f;
    if (idx < 16 * 33 && b < B) {
      if (j < 32) {
        ex[i] = __ldg(s.dposeA + (size_t)b * 320 + r * 32 + j);
        if (a.route_mask) rmv[i] = __ldg(a.route_mask + (size_t)b * 10 + r);
      } else if (!a.d.from_poses) {
        ex[i] = __ldg(a.dpc + (size_t)b * 330 + r * 33 + 32);      // written by rs_iterate_bwd_kernel
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int idx = tid + 256 * i, p = idx / 33, j = idx % 33, b = b0 + p;
    if (idx >= 16 * 33) break;
    float v = 0.f;
    if (b < B) {
      if (j < 32) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += part[w][p][j];
        v = (sum * (1.0f / RS_DU_SCALE) + ex[i]) * rmv[i];
        if (a.d.from_poses) { if (a.d_poses) a.d_poses[(size_t)b * 320 + r * 32 + j] = v; }
        else a.dpc[(size_t)b * 330 + r * 33 + j] = v;
      } else {
        v = ex[i];
      }
    }
    dps[p][j] = v;
  }
  if (a.d.from_poses) return;
  __syncthreads();
  if (tid < 33 && a.d_proj_b[r]) {      // projector bias gradient of this (tile, route)
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < 16; ++p) sum += dps[p][tid];
    atomicAdd(a.d_proj_b[r] + tid, sum);
  }
  if (!a.d_route_embs) return;
  // d e[b][r][c] = sum_j dpc[b][r][j] W_r[j][c]  as [16 patients x 40 (33, zero padded)] . [40 x 256] on mma.sync m16n8k8 tf32
  // (the gradients keep their exponent range in tf32; the CUDA-core form was 16 x 33 FMAs per thread and, at K = 2 / B = 8192,
  // the largest item of the routing backward).  Warp w owns columns [32 w, 32 w + 32).
  {
    const float* W = a.p.proj_w[r];
    float c[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 5; ++ks) {
      const int j0 = 8 * ks + t, j1 = j0 + 4;
      const uint32_t a0 = to_tf32(dps[g][j0]), a1 = to_tf32(dps[g + 8][j0]);
      const uint32_t a2 = to_tf32(dps[g][j1]), a3 = to_tf32(dps[g + 8][j1]);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int col = 32 * warp + 8 * n + g;
        const uint32_t b0 = j0 < 33 ? to_tf32(__ldg(W + (size_t)j0 * 256 + col)) : 0u;
        const uint32_t b1 = j1 < 33 ? to_tf32(__ldg(W + (size_t)j1 * 256 + col)) : 0u;
        mma_tf32(c[n], a0, a1, a2, a3, b0, b1);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = 32 * warp + 8 * n + 2 * t;
      if (vl) *reinterpret_cast<float2*>(a.d_route_embs + (size_t)r * a.d.emb_route_stride + (size_t)(b0 + g) * a.d.emb_batch_stride + col) = make_float2(c[n][0], c[n][1]);
      if (vh) *reinterpret_cast<float2*>(a.d_route_embs + (size_t)r * a.d.emb_route_stride + (size_t)(b0 + g + 8) * a.d.emb_batch_stride + col) = make_float2(c[n][2], c[n][3]);
    }
  }
}

}  // namespace mmr
