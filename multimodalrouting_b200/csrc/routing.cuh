// Capsule-style routing + heads.  One CTA owns a tile of PB patients and keeps ALL of their routing
// state in shared memory across the agreement iterations: the 10 route embeddings, the primary
// capsules, the 10 x K x 64 votes (fp32 in the fp32 parity mode, saturating fp16 in the bf16 mode -- the
// reference computes these einsums in bf16 under autocast), per-iteration decision poses and routing
// coefficients.  The vote weights w[10,32,K,64] (2 MB at K=25) and the projector weights stream
// through the CTA ONCE per tile (coalesced 16-byte loads, PB FMAs per loaded value), so their L2
// traffic is amortised over the tile instead of being re-read per patient.
//   RoutePrimaryProjector.forward         routing_and_heads.py:111-121
//   forward_capsule_from_route_dict       routing_and_heads.py:314-352 (mask / temperature / clamp)
//   CapsuleMortalityHead.forward          Mort :194-268 / Pheno :194-272
//   CapsuleFC.forward                     capsule_layers.py:75-117
#pragma once
#include <type_traits>

#include "mmr_common.cuh"

namespace mmr {

#ifndef MMR_RT_THREADS
#define MMR_RT_THREADS 512
#endif
#ifndef MMR_RT_MINB
#define MMR_RT_MINB 1
#endif
constexpr int RT_THREADS = MMR_RT_THREADS;    // 16 warps: one CTA per SM, latency hidden by warps rather than by CTAs
constexpr int RT_GSLOTS = 1024 / RT_THREADS;   // K*32 <= 1024 head-gradient accumulators spread over the CTA
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
  float* d_proj_b[MMR_ROUTES];   // split path: projector bias gradients (accumulated by rs_dpose_kernel) or null
};

// per-patient persistent state (floats)
constexpr int RT_PP_FWD = 320 + 7 * 16;            // pose, zl, a0, a2, a3, alpha, act, rm
constexpr int RT_PP_BWD = RT_PP_FWD + 320 + 48;    // + dpose, misc (d alpha | d act | d act-logit)

struct RtPatient {   // views into the per-patient block
  float* pose; float* zl; float* a0; float* a2; float* a3; float* alpha; float* act; float* rm;
  float* dpose; float* misc;
};
__device__ __forceinline__ RtPatient rt_patient(float* base, int p, bool bwd) {
  float* q = base + (size_t)p * (bwd ? RT_PP_BWD : RT_PP_FWD);
  RtPatient s;
  s.pose = q; q += 320;
  s.zl = q; q += 16; s.a0 = q; q += 16; s.a2 = q; q += 16; s.a3 = q; q += 16;
  s.alpha = q; q += 16; s.act = q; q += 16; s.rm = q; q += 16;
  s.dpose = q; s.misc = q + 320;
  return s;
}

// scratch shared by the patients of a tile (one patient walks the iterations at a time)
struct RtScratch {
  float* G;      // [K][32]
  float* v;      // [nit][K*64]
  float* q;      // [nit][10*K]
  float* Rn;     // [10*K]
  float* dp;     // [K][32] decision poses d_bkp
  float* dv;     // bwd [nit][K*64]
  float* ds;     // bwd [nit][10*K]
  float* ddp;    // bwd [K][32]
  float* tr;     // bwd [10*K]
};
__host__ __device__ inline size_t rt_scratch_floats(int K, int nit, bool bwd) {
  size_t n = (size_t)K * 32 + (size_t)nit * K * 64 + (size_t)nit * 10 * K + 10 * K + K * 32;
  if (bwd) n += (size_t)nit * K * 64 + (size_t)nit * 10 * K + K * 32 + 10 * K;
  return (n + 3) / 4 * 4;
}
// vote region per patient: max(votes in UT, the 10 route embeddings in fp32 that alias it during the projector phase)
__host__ __device__ inline size_t rt_u_bytes(int K, size_t ut_size, bool from_poses) {
  const size_t u = (size_t)10 * K * 64 * ut_size, e = from_poses ? 0 : (size_t)10 * 256 * 4;
  return ((u > e ? u : e) + 15) / 16 * 16;
}
__host__ __device__ inline size_t rt_smem_bytes(int K, int nit, bool bwd, int PB, size_t ut_size, bool from_poses) {
  return rt_scratch_floats(K, nit, bwd) * 4 + (size_t)PB * (bwd ? RT_PP_BWD : RT_PP_FWD) * 4 +
         (size_t)PB * rt_u_bytes(K, ut_size, from_poses) + 16;
}
__device__ inline RtScratch rt_carve(float* base, int K, int nit, bool bwd) {
  RtScratch s;
  float* p = base;
  s.G = p; p += K * 32;
  s.v = p; p += nit * K * 64;
  s.q = p; p += nit * 10 * K;
  s.Rn = p; p += 10 * K;
  s.dp = p; p += K * 32;
  if (bwd) {
    s.dv = p; p += nit * K * 64;
    s.ds = p; p += nit * 10 * K;
    s.ddp = p; p += K * 32;
    s.tr = p; p += 10 * K;
  } else {
    s.dv = s.ds = s.ddp = s.tr = nullptr;
  }
  return s;
}

// G[k][p] = sum_m pose_to_mc[m][p] * embedding[k][m]   (logits = sum_p d[k][p]*G[k][p] + bias[k])
__device__ inline void rt_build_G(const RoutingArgs& a, float* G) {
  const int K = a.d.K;
  for (int i = threadIdx.x; i < K * 32; i += RT_THREADS) {
    const int k = i >> 5, p = i & 31;
    float acc = 0.f;
    for (int m = 0; m < MC; ++m) acc = fmaf(a.p.pose_to_mc[m * PC + p], a.p.embedding[k * MC + m], acc);
    G[i] = acc;
  }
}

// ---- tensor-core forms of the projector and the vote contraction (reduced-precision mode) ------------------
// Both are small GEMMs whose M dimension is the patient tile (PB <= 8 rows of an m16n8k16 tile; rows >= PB are zero)
// and whose B operand is a weight matrix pre-packed to fp16 with the reduction index contiguous
// (mmr_routing_pack_weights): caps_wt[r][c][a] (a = 32) and proj_wb[r][n (40, rows >= 33 zero)][k (256)].
// The reduction index is permuted so that ONE 16-byte load per lane yields the B fragments of two k-steps:
// logical k-slot (step s, column 2t+e / 2t+8+e of the fragment) <-> physical index 8t + 4s + e / 8t + 4s + 2 + e
// inside a group of 32; the A fragments are built from shared memory with the same permutation.
// operands are fp16 (11-bit mantissa, saturating), like the votes held in shared memory: 8x tighter than the bf16 the
// reference's autocast einsums use -- sharpened routing amplifies operand rounding (tests/golden pheno_sharp4)
__device__ __forceinline__ uint32_t rt_pack2(float lo, float hi) {
  const float L = 65504.f;
  __half2 v = __floats2half2_rn(fminf(fmaxf(lo, -L), L), fminf(fmaxf(hi, -L), L));
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void rt_mma(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  const uint32_t z = 0u;    // rows 8..15 of the A tile are unused
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(z), "r"(a2), "r"(z), "r"(b0), "r"(b1));
}
// A fragments (two k-steps) of 8 consecutive fp32 values x[8t .. 8t+7] of this lane's row
__device__ __forceinline__ void rt_afrag(const float* x8, bool valid, uint32_t (&A)[4]) {
  A[0] = A[1] = A[2] = A[3] = 0u;
  if (valid) {
    const float4 x0 = *reinterpret_cast<const float4*>(x8), x1 = *reinterpret_cast<const float4*>(x8 + 4);
    A[0] = rt_pack2(x0.x, x0.y); A[1] = rt_pack2(x0.z, x0.w); A[2] = rt_pack2(x1.x, x1.y); A[3] = rt_pack2(x1.z, x1.w);
  }
}

// projector: pose[p][r][j] / zl[p][r] = sum_k emb[p][r][k] * W_r[j][k] + b_r[j]   (one warp per route)
template <int PB>
__device__ inline void rt_project_mma(const RoutingArgs& a, float* pp, const uint8_t* ureg, size_t ustride, bool bwd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint4* pw = reinterpret_cast<const uint4*>(a.p.proj_w_f16);
  for (int r = warp; r < 10; r += RT_THREADS / 32) {
    float acc[5][4];
#pragma unroll
    for (int n = 0; n < 5; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    const float* e = reinterpret_cast<const float*>(ureg + (g < PB ? g : 0) * ustride) + r * 256 + 8 * t;
#pragma unroll 2
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t A[4];
      rt_afrag(e + 32 * kk, g < PB, A);
#pragma unroll
      for (int n = 0; n < 5; ++n) {
        const uint4 q = __ldg(pw + ((size_t)(r * 40 + n * 8 + g) * 32 + kk * 4 + t));
        rt_mma(acc[n], A[0], A[1], q.x, q.y);
        rt_mma(acc[n], A[2], A[3], q.z, q.w);
      }
    }
    if (g < PB) {
      RtPatient s = rt_patient(pp, g, bwd);
#pragma unroll
      for (int n = 0; n < 5; ++n)
#pragma unroll
        for (int ee = 0; ee < 2; ++ee) {
          const int j = n * 8 + 2 * t + ee;
          if (j < 33) {
            const float v = acc[n][ee] + a.p.proj_b[r][j];
            if (j < 32) s.pose[r * 32 + j] = v; else s.zl[r] = v;
          }
        }
    }
  }
}

// votes: u[p][r][c] = sum_a pose[p][r][a] * w[r][a][c]   (n-tiles of 8 columns round-robin over the warps)
template <int PB>
__device__ inline void rt_votes_mma(const RoutingArgs& a, const float* pp, uint8_t* ureg, size_t ustride, bool bwd) {
  const int KD = a.d.K * 64, NT = KD / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int ppstride = bwd ? RT_PP_BWD : RT_PP_FWD;
  constexpr int NW = RT_THREADS / 32;
  __half* urow = reinterpret_cast<__half*>(ureg + (g < PB ? g : 0) * ustride);
  for (int r = 0; r < 10; ++r) {
    uint32_t A[4];
    rt_afrag(pp + (g < PB ? g : 0) * ppstride + r * 32 + 8 * t, g < PB, A);
    const uint4* wr = reinterpret_cast<const uint4*>(a.p.caps_wt_f16) + (size_t)r * KD * 4 + (size_t)g * 4 + t;
    for (int nt0 = warp; nt0 < NT; nt0 += 4 * NW) {
      uint4 q[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nt = nt0 + i * NW;
        q[i] = nt < NT ? __ldg(wr + (size_t)nt * 32) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int nt = nt0 + i * NW;
        if (nt >= NT) break;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        rt_mma(acc, A[0], A[1], q[i].x, q[i].y);
        rt_mma(acc, A[2], A[3], q[i].z, q[i].w);
        if (g < PB) {
          const float L = 65504.f;
          *reinterpret_cast<__half2*>(urow + r * KD + nt * 8 + 2 * t) =
              __floats2half2_rn(fminf(fmaxf(acc[0], -L), L), fminf(fmaxf(acc[1], -L), L));
        }
      }
    }
  }
  __syncthreads();
}

// backward of the vote contraction: dpose[p][r][a] += sum_c du[p][r][c] * w[r][a][c]  (du already fp16 in the vote region;
// (route, 32-column group) units are split evenly over the warps, partial tiles are flushed with shared-memory atomics)
template <int PB>
__device__ inline void rt_dpose_mma(const RoutingArgs& a, float* pp, const uint8_t* ureg, size_t ustride, int np) {
  const int KD = a.d.K * 64, G = KD / 32, U = 10 * G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  constexpr int NW = RT_THREADS / 32;
  const int u0 = (int)((long long)U * warp / NW), u1 = (int)((long long)U * (warp + 1) / NW);
  const __half* urow = reinterpret_cast<const __half*>(ureg + (g < PB ? g : 0) * ustride);
  const uint4* w16 = reinterpret_cast<const uint4*>(a.p.caps_w_f16);
  float acc[4][4];
  int rcur = -1;
  auto flush = [&]() {
    if (rcur < 0 || g >= PB || g >= np) return;
    float* dp = rt_patient(pp, g, true).dpose + rcur * 32;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      atomicAdd(dp + n * 8 + 2 * t, acc[n][0]);
      atomicAdd(dp + n * 8 + 2 * t + 1, acc[n][1]);
    }
  };
  for (int u = u0; u < u1; ++u) {
    const int r = u / G, kk = u % G;
    if (r != rcur) {
      flush();
      rcur = r;
#pragma unroll
      for (int n = 0; n < 4; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
    }
    uint4 A = make_uint4(0u, 0u, 0u, 0u);
    if (g < PB) A = *reinterpret_cast<const uint4*>(urow + r * KD + 32 * kk + 8 * t);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const uint4 q = __ldg(w16 + ((size_t)(r * 32 + n * 8 + g) * KD + 32 * kk + 8 * t) / 8);
      rt_mma(acc[n], A.x, A.y, q.x, q.y);
      rt_mma(acc[n], A.z, A.w, q.z, q.w);
    }
  }
  flush();
}

// ---- tile phase 1: projector + activation chain for the np patients b0 .. b0+np-1 ---------------
template <int PB>
__device__ inline void rt_project(const RoutingArgs& a, float* pp, uint8_t* ureg, size_t ustride, bool bwd, int b0, int np) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_mask = a.route_mask != nullptr;
  for (int i = tid; i < PB * 10; i += RT_THREADS) {
    const int p = i / 10, r = i % 10;
    RtPatient s = rt_patient(pp, p, bwd);
    s.rm[r] = (p < np && has_mask) ? a.route_mask[(size_t)(b0 + p) * 10 + r] : 1.f;
  }
  if (!a.d.from_poses) {
    // route embeddings -> the (not yet used) vote region
    for (int i = tid; i < PB * 10 * 64; i += RT_THREADS) {
      const int p = i / 640, r = (i % 640) >> 6, c = (i & 63) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < np)
        v = *reinterpret_cast<const float4*>(a.route_embs + (size_t)r * a.d.emb_route_stride +
                                             (size_t)(b0 + p) * a.d.emb_batch_stride + c);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(ureg + p * ustride) + r * 256 + c) = v;
    }
    __syncthreads();
    if (a.p.proj_w_f16 != nullptr && a.d.vote_dtype == MMR_DTYPE_BF16 && PB <= 8) {
      rt_project_mma<PB>(a, pp, ureg, ustride, bwd);   // tensor cores, fp16 operands, fp32 accumulation
    } else {
      // 10 x 33 dot products of length 256: one warp per output, the weight row is shared by the tile
      for (int o = warp; o < 330; o += RT_THREADS / 32) {
        const int r = o / 33, j = o % 33;
        const float* w = a.p.proj_w[r] + (size_t)j * 256;
        float wv[8];
  #pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = __ldg(w + lane + 32 * i);
        const float bias = a.p.proj_b[r][j];
  #pragma unroll
        for (int p = 0; p < PB; ++p) {
          const float* e = reinterpret_cast<const float*>(ureg + p * ustride) + r * 256;
          float acc = 0.f;
  #pragma unroll
          for (int i = 0; i < 8; ++i) acc = fmaf(wv[i], e[lane + 32 * i], acc);
          acc = warp_sum(acc);
          if (lane == 0) {
            RtPatient s = rt_patient(pp, p, bwd);
            if (j < 32) s.pose[r * 32 + j] = acc + bias; else s.zl[r] = acc + bias;
          }
        }
      }
    }
    __syncthreads();
    for (int i = tid; i < PB * 10; i += RT_THREADS) {
      const int p = i / 10, r = i % 10;
      if (p >= np) continue;
      RtPatient s = rt_patient(pp, p, bwd);
      const int b = b0 + p;
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
      for (int i = tid; i < np * 320; i += RT_THREADS) {
        const int p = i / 320, j = i % 320;
        a.poses_out[(size_t)(b0 + p) * 320 + j] = rt_patient(pp, p, bwd).pose[j];
      }
    __syncthreads();
  } else {
    for (int i = tid; i < PB * 320; i += RT_THREADS) {
      const int p = i / 320, j = i % 320;
      rt_patient(pp, p, bwd).pose[j] = p < np ? a.poses_in[(size_t)(b0 + p) * 320 + j] : 0.f;
    }
    for (int i = tid; i < PB * 10; i += RT_THREADS) {
      const int p = i / 10, r = i % 10;
      RtPatient s = rt_patient(pp, p, bwd);
      const float x = p < np ? a.acts_in[(size_t)(b0 + p) * 10 + r] : 0.f;
      s.a0[r] = s.a2[r] = s.a3[r] = x;
      s.alpha[r] = has_mask ? x * s.rm[r] : x;
    }
    __syncthreads();
  }
  // mask poses, pick the routing activation
  for (int i = tid; i < PB * 320; i += RT_THREADS) {
    const int p = i / 320, j = i % 320;
    RtPatient s = rt_patient(pp, p, bwd);
    s.pose[j] *= s.rm[j >> 5];
  }
  for (int i = tid; i < PB * 10; i += RT_THREADS) {
    const int p = i / 10, r = i % 10;
    RtPatient s = rt_patient(pp, p, bwd);
    s.act[r] = (a.d.variant == MMR_VARIANT_PHENO) ? s.alpha[r] : s.rm[r];
  }
  __syncthreads();
}

// ---- tile phase 2: votes u[p][r][c] = sum_a pose[p][r][a] * w[r][a][c],  c = k*64 + d ------------
// A thread owns 4 consecutive columns; every 16-byte weight load feeds PB*4 FMAs.
template <int PB, class UT>
__device__ inline void rt_votes(const RoutingArgs& a, const float* pp, uint8_t* ureg, size_t ustride, bool bwd) {
  if constexpr (std::is_same<UT, __half>::value && PB <= 8) {
    if (a.p.caps_wt_f16 != nullptr) {
      rt_votes_mma<PB>(a, pp, ureg, ustride, bwd);
      return;
    }
  }
  const int KD = a.d.K * 64;
  const int ppstride = bwd ? RT_PP_BWD : RT_PP_FWD;
  for (int cg = threadIdx.x; cg < KD / 4; cg += RT_THREADS) {
    for (int r = 0; r < 10; ++r) {
      float4 acc[PB];
#pragma unroll
      for (int p = 0; p < PB; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* w = reinterpret_cast<const float4*>(a.p.caps_w + (size_t)r * 32 * KD) + cg;
#pragma unroll 2
      for (int a4 = 0; a4 < 32; a4 += 4) {
        float4 w4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w4[j] = __ldg(w + (size_t)(a4 + j) * (KD / 4));
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          const float4 x = *reinterpret_cast<const float4*>(pp + p * ppstride + r * 32 + a4);   // broadcast
          const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[p].x = fmaf(xs[j], w4[j].x, acc[p].x); acc[p].y = fmaf(xs[j], w4[j].y, acc[p].y);
            acc[p].z = fmaf(xs[j], w4[j].z, acc[p].z); acc[p].w = fmaf(xs[j], w4[j].w, acc[p].w);
          }
        }
      }
#pragma unroll
      for (int p = 0; p < PB; ++p) Vec4<UT>::st(reinterpret_cast<UT*>(ureg + p * ustride) + r * KD + 4 * cg, acc[p]);
    }
  }
  __syncthreads();
}

// dot_d u[k*64 + d] * v[k*64 + d] with the d index rotated by 2k so that the lanes of a warp (consecutive k)
// hit distinct shared-memory banks although their rows are 64 elements apart.
template <class UT> __device__ __forceinline__ float rt_dot64(const UT* urow, const float* vrow, int k);
template <> __device__ __forceinline__ float rt_dot64<float>(const float* urow, const float* vrow, int k) {
  float acc = 0.f;
#pragma unroll 8
  for (int d = 0; d < 64; d += 2) {
    const int dd = (d + 2 * k) & 63;
    const float2 a = *reinterpret_cast<const float2*>(urow + dd), b = *reinterpret_cast<const float2*>(vrow + dd);
    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
  }
  return acc;
}
template <> __device__ __forceinline__ float rt_dot64<__half>(const __half* urow, const float* vrow, int k) {
  float acc = 0.f;
#pragma unroll 8
  for (int d = 0; d < 64; d += 2) {
    const int dd = (d + 2 * k) & 63;
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(urow + dd));
    const float2 b = *reinterpret_cast<const float2*>(vrow + dd);
    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
  }
  return acc;
}

// ---- per-patient agreement iterations; leaves q / v / Rn / dp in the shared scratch -------------
template <class UT>
__device__ inline void rt_iterate(const RoutingArgs& a, const RtScratch& s, const RtPatient& pt, const UT* u) {
  const int K = a.d.K, KD = K * 64, nit = a.d.num_routing;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float scale = 0.125f;   // 1/sqrt(64)
  const float invK = 1.0f / (float)K;
  for (int it = 0; it < nit; ++it) {
    float* q = s.q + it * 10 * K;
    if (it == 0) {
      for (int i = tid; i < 10 * K; i += RT_THREADS) q[i] = invK;
    } else {
      const float* vp = s.v + (it - 1) * KD;
      for (int o = tid; o < 10 * K; o += RT_THREADS) {      // one thread per (route, label) agreement
        const int r = o / K, k = o % K;
        q[o] = scale * rt_dot64<UT>(u + r * KD + k * 64, vp + k * 64, k);
      }
      __syncthreads();
      for (int r = warp; r < 10; r += RT_THREADS / 32) {    // softmax over labels: one warp per route (K <= 32)
        const float x = lane < K ? q[r * K + lane] : -INFINITY;
        const float mx = warp_max(x);
        const float ex = lane < K ? expf(x - mx) : 0.f;
        const float pr = ex * (1.0f / warp_sum(ex));
        const float t = warp_sum(pr);
        if (lane < K) q[r * K + lane] = pr * (1.0f / (t + 1e-10f));
      }
    }
    __syncthreads();
    if (it + 1 < nit) {   // the decision pose of the last iteration is never consumed
      float* vo = s.v + it * KD;
      for (int c = tid; c < KD; c += RT_THREADS) {
        const int k = c >> 6;
        float acc = 0.f;
        if (it == 0) {
          for (int r = 0; r < 10; ++r) acc += to_f<UT>(u[r * KD + c]);
          acc *= invK;
        } else {
          for (int r = 0; r < 10; ++r) acc = fmaf(q[r * K + k] * pt.act[r], to_f<UT>(u[r * KD + c]), acc);
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
    for (int r = 0; r < 10; ++r) den += ql[r * K + tid] * pt.rm[r];
    den = fmaxf(den, 1e-10f);
    for (int r = 0; r < 10; ++r) s.Rn[r * K + tid] = ql[r * K + tid] * pt.rm[r] / den;
  }
  __syncthreads();
  for (int i = tid; i < K * 32; i += RT_THREADS) {
    const int k = i >> 5, p = i & 31;
    float acc = 0.f;
    for (int r = 0; r < 10; ++r) {
      const float cr = (a.d.variant == MMR_VARIANT_PHENO) ? pt.alpha[r] : 1.f;
      acc = fmaf(s.Rn[r * K + k] * cr, pt.pose[r * 32 + p], acc);
    }
    s.dp[i] = acc;
  }
  __syncthreads();
}

template <int PB, class UT>
__global__ void __launch_bounds__(RT_THREADS, MMR_RT_MINB) routing_fwd_kernel(RoutingArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = a.d.K, nit = a.d.num_routing;
  float* sf = reinterpret_cast<float*>(smem_raw);
  RtScratch s = rt_carve(sf, K, nit, false);
  float* pp = sf + rt_scratch_floats(K, nit, false);
  uint8_t* ureg = reinterpret_cast<uint8_t*>(pp + PB * RT_PP_FWD);
  const size_t ustride = rt_u_bytes(K, sizeof(UT), a.d.from_poses != 0);
  rt_build_G(a, s.G);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (a.d.B + PB - 1) / PB;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * PB, np = min(PB, a.d.B - b0);
    __syncthreads();
    rt_project<PB>(a, pp, ureg, ustride, false, b0, np);
    rt_votes<PB, UT>(a, pp, ureg, ustride, false);
    for (int p = 0; p < np; ++p) {
      const int b = b0 + p;
      RtPatient pt = rt_patient(pp, p, false);
      rt_iterate<UT>(a, s, pt, reinterpret_cast<const UT*>(ureg + p * ustride));
      for (int k = warp; k < K; k += RT_THREADS / 32) {
        float acc = s.dp[k * 32 + lane] * s.G[k * 32 + lane];
        acc = warp_sum(acc);
        if (lane == 0) a.logits[(size_t)b * K + k] = acc + a.p.bias[k];
      }
      if (tid < 10) a.alpha[(size_t)b * 10 + tid] = pt.alpha[tid];
      if (a.R)
        for (int i = tid; i < 10 * K; i += RT_THREADS) a.R[(size_t)b * 10 * K + i] = s.Rn[i];
      __syncthreads();
    }
  }
}

template <int PB, class UT>
__global__ void __launch_bounds__(RT_THREADS, MMR_RT_MINB) routing_bwd_kernel(RoutingArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int K = a.d.K, KD = K * 64, nit = a.d.num_routing;
  float* sf = reinterpret_cast<float*>(smem_raw);
  RtScratch s = rt_carve(sf, K, nit, true);
  float* pp = sf + rt_scratch_floats(K, nit, true);
  uint8_t* ureg = reinterpret_cast<uint8_t*>(pp + PB * RT_PP_BWD);
  const size_t ustride = rt_u_bytes(K, sizeof(UT), a.d.from_poses != 0);
  rt_build_G(a, s.G);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool pheno = a.d.variant == MMR_VARIANT_PHENO;
  const bool has_mask = a.route_mask != nullptr;
  const float scale = 0.125f, invK = 1.0f / (float)K;
  // per-CTA accumulators for dG / dbias (flushed once at the end)
  float accG[RT_GSLOTS];                    // element i = tid + RT_THREADS*j of [K][32]  (K*32 <= 1024)
#pragma unroll
  for (int j = 0; j < RT_GSLOTS; ++j) accG[j] = 0.f;
  float accB = 0.f;                         // tid < K
  const int ntiles = (a.d.B + PB - 1) / PB;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * PB, np = min(PB, a.d.B - b0);
    __syncthreads();
    rt_project<PB>(a, pp, ureg, ustride, true, b0, np);
    rt_votes<PB, UT>(a, pp, ureg, ustride, true);
    for (int p = 0; p < np; ++p) {
      const int b = b0 + p;
      RtPatient pt = rt_patient(pp, p, true);
      UT* u = reinterpret_cast<UT*>(ureg + p * ustride);
      rt_iterate<UT>(a, s, pt, u);
      const float* dl = a.d_logits + (size_t)b * K;
      // (a,b) head: dG, dbias, ddp = dlogit[k]*G[k][p]
#pragma unroll
      for (int j = 0; j < RT_GSLOTS; ++j) {
        const int i = tid + RT_THREADS * j;
        if (i < K * 32) {
          const float g = dl[i >> 5];
          accG[j] = fmaf(g, s.dp[i], accG[j]);
          s.ddp[i] = g * s.G[i];
        }
      }
      if (tid < K) accB += dl[tid];
      for (int i = tid; i < 320; i += RT_THREADS) a.poses_m[(size_t)b * 320 + i] = pt.pose[i];
      __syncthreads();
      // (c) t_rk = sum_p ddp[k][p]*pose[r][p];  dRt = dR + c_r*t_rk  (held in ds[nit-1])
      float* dRt = s.ds + (nit - 1) * 10 * K;
      for (int o = tid; o < 10 * K; o += RT_THREADS) {
        const int r = o / K, k = o % K;
        float t = 0.f;
#pragma unroll 8
        for (int c = 0; c < 32; ++c) {      // rotated by k: conflict-free although the ddp rows are 32 words apart
          const int cc = (c + k) & 31;
          t = fmaf(s.ddp[k * 32 + cc], pt.pose[r * 32 + cc], t);
        }
        const float cr = pheno ? pt.alpha[r] : 1.f;
        const float dr = a.d_R ? a.d_R[(size_t)b * 10 * K + o] : 0.f;
        s.tr[o] = t;
        dRt[o] = dr + cr * t;
      }
      __syncthreads();
      // dpose from the final aggregation, dalpha (pheno)
      for (int i = tid; i < 320; i += RT_THREADS) {
        const int r = i >> 5, c = i & 31;
        const float cr = pheno ? pt.alpha[r] : 1.f;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(s.Rn[r * K + k], s.ddp[k * 32 + c], acc);
        pt.dpose[i] = cr * acc;
      }
      if (tid < 10) {
        float da = 0.f;
        if (pheno)
          for (int k = 0; k < K; ++k) da = fmaf(s.Rn[tid * K + k], s.tr[tid * K + k], da);
        pt.misc[tid] = da;          // d alpha
        pt.misc[16 + tid] = 0.f;    // d act (filled by the iterations)
      }
      __syncthreads();
      // (d) through R = qm / den
      const float* ql = s.q + (nit - 1) * 10 * K;
      if (tid < K) {
        const int k = tid;
        float den = 0.f, dot = 0.f;
        for (int r = 0; r < 10; ++r) { den += ql[r * K + k] * pt.rm[r]; dot = fmaf(dRt[r * K + k], s.Rn[r * K + k], dot); }
        const bool clamped = den < 1e-10f;
        const float dd = clamped ? 1e-10f : den;
        for (int r = 0; r < 10; ++r) {
          const float g = clamped ? dRt[r * K + k] : (dRt[r * K + k] - dot);
          dRt[r * K + k] = g / dd * pt.rm[r];     // now holds dq of the last iteration
        }
      }
      __syncthreads();
      // (e) agreement iterations, last to first.  ds[it] holds dq_it on entry and ds_it on exit.
      for (int it = nit - 1; it >= 1; --it) {
        float* dq = s.ds + it * 10 * K;
        const float* q = s.q + it * 10 * K;
        for (int r = warp; r < 10; r += RT_THREADS / 32) {   // one warp per route, lane = label
          const float g = lane < K ? dq[r * K + lane] : 0.f, qq = lane < K ? q[r * K + lane] : 0.f;
          const float aa = warp_sum(g * qq);
          const float bb = warp_sum((g - aa) * qq);   // renormalisation by (sum softmax + 1e-10) == 1 up to rounding
          if (lane < K) dq[r * K + lane] = qq * ((g - aa) - bb);
        }
        __syncthreads();
        // dv_{it-1}[c] = scale * sum_r ds[r][k] * u[r][c]
        float* dvp = s.dv + (it - 1) * KD;
        for (int c = tid; c < KD; c += RT_THREADS) {
          const int k = c >> 6;
          float acc = 0.f;
          for (int r = 0; r < 10; ++r) acc = fmaf(dq[r * K + k], to_f<UT>(u[r * KD + c]), acc);
          dvp[c] = acc * scale;
        }
        __syncthreads();
        if (it - 1 >= 1) {
          // v_{it-1} = sum_r q_{it-1}*act*u  ->  dq_{it-1}[r][k] = act[r]*w_rk, dact[r] += sum_k q*w_rk
          float* dqp = s.ds + (it - 1) * 10 * K;
          const float* qp = s.q + (it - 1) * 10 * K;
          for (int o = tid; o < 10 * K; o += RT_THREADS) {
            const int r = o / K, k = o % K;
            const float w = rt_dot64<UT>(u + r * KD + k * 64, dvp + k * 64, k);
            dqp[o] = pt.act[r] * w;
            atomicAdd(&pt.misc[16 + r], qp[o] * w);
          }
          __syncthreads();
        }
      }
      // (f) du[r][c], written to global (fp32, operand of the vote-weight gradient) and in place of u
      for (int c = tid; c < KD; c += RT_THREADS) {
        const int k = c >> 6;
        for (int r = 0; r < 10; ++r) {
          float du = 0.f;
          if (nit >= 2) du = invK * s.dv[c];
          for (int it = 1; it < nit; ++it) {
            du = fmaf(s.ds[it * 10 * K + r * K + k], scale * s.v[(it - 1) * KD + c], du);
            if (it < nit - 1) du = fmaf(s.q[it * 10 * K + r * K + k] * pt.act[r], s.dv[it * KD + c], du);
          }
          a.du[((size_t)b * 10 + r) * KD + c] = du;
          u[r * KD + c] = from_f<UT>(du);
        }
      }
      __syncthreads();
    }
    // (f') dpose[p][r][a] += sum_c du[p][r][c]*w[r][a][c]: a thread owns (column quarter, r, a); lanes of a
    // warp share r, so the du reads are shared-memory broadcasts and each weight load feeds PB dot products.
    bool dpose_done = false;
    if constexpr (std::is_same<UT, __half>::value && PB <= 8) {
      if (a.p.caps_w_f16 != nullptr) {
        rt_dpose_mma<PB>(a, pp, ureg, ustride, np);
        dpose_done = true;
      }
    }
    if (!dpose_done)
    for (int item = tid; item < 4 * 320; item += RT_THREADS) {
      const int cq = item / 320, ra = item % 320, r = ra >> 5;
      const int c4_lo = (KD / 4) * cq / 4, c4_hi = (KD / 4) * (cq + 1) / 4;
      const float4* w = reinterpret_cast<const float4*>(a.p.caps_w + (size_t)ra * KD);
      float acc[PB];
#pragma unroll
      for (int p = 0; p < PB; ++p) acc[p] = 0.f;
#pragma unroll 4
      for (int c4 = c4_lo; c4 < c4_hi; ++c4) {
        const float4 w4 = __ldg(w + c4);
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          const float4 d4 = Vec4<UT>::ld(reinterpret_cast<const UT*>(ureg + p * ustride) + r * KD + 4 * c4);
          acc[p] = fmaf(d4.x, w4.x, fmaf(d4.y, w4.y, fmaf(d4.z, w4.z, fmaf(d4.w, w4.w, acc[p]))));
        }
      }
#pragma unroll
      for (int p = 0; p < PB; ++p)
        if (p < np) atomicAdd(&rt_patient(pp, p, true).dpose[ra], acc[p]);
    }
    __syncthreads();
    // (g) activation chain and projector data-gradient
    for (int i = tid; i < np * 10; i += RT_THREADS) {
      const int p = i / 10, r = i % 10, b = b0 + p;
      RtPatient pt = rt_patient(pp, p, true);
      const float rm = pt.rm[r];
      float dal = pt.misc[r] + (pheno ? pt.misc[16 + r] : 0.f);   // d alpha (alpha = a3*rm)
      float g = has_mask ? dal * rm : dal;                         // d a3
      if (a.d.from_poses) {
        if (a.d_acts) a.d_acts[(size_t)b * 10 + r] = g;
        pt.misc[32 + r] = 0.f;
      } else {
        if (a.d.detach_priors) g = 0.f;
        const bool keep = has_mask ? (rm != 0.f) : true;
        if (keep) { const float x = pt.a2[r]; if (x < a.d.prior_floor || x > a.d.prior_ceiling) g = 0.f; }
        if (a.d.act_temperature != 1.0f && has_mask && keep) {
          const float a1 = pt.a0[r] * rm;
          if (a1 < 1e-6f || a1 > 1.0f - 1e-6f) g = 0.f;
          else { const float y = pt.a2[r]; g *= y * (1.0f - y) / a.d.act_temperature / (a1 * (1.0f - a1)); }
        }
        if (has_mask) g *= rm;
        if (a.acts_override) {     // the chain ends at the externally supplied prior (routing_and_heads.py:314)
          if (a.d_acts) a.d_acts[(size_t)b * 10 + r] = g;
          g = 0.f;
        } else {
          g *= pt.a0[r] * (1.0f - pt.a0[r]);
        }
        pt.misc[32 + r] = g;   // d (activation logit)
      }
    }
    __syncthreads();
    if (a.d.from_poses) {
      if (a.d_poses)
        for (int i = tid; i < np * 320; i += RT_THREADS) {
          const int p = i / 320, j = i % 320;
          RtPatient pt = rt_patient(pp, p, true);
          a.d_poses[(size_t)(b0 + p) * 320 + j] = pt.dpose[j] * pt.rm[j >> 5];
        }
    } else {
      // dpc[p][r][j] (j < 32: pose, j = 32: activation logit); staged in the pose slots (no longer needed)
      for (int i = tid; i < np * 330; i += RT_THREADS) {
        const int p = i / 330, o = i % 330, r = o / 33, j = o % 33;
        RtPatient pt = rt_patient(pp, p, true);
        const float g = (j < 32) ? pt.dpose[r * 32 + j] * pt.rm[r] : pt.misc[32 + r];
        a.dpc[(size_t)(b0 + p) * 330 + o] = g;
      }
      __syncthreads();
      if (a.d_route_embs) {
        // d e[p][r][c] = sum_j dpc[p][r][j] * W_r[j][c]; thread = column c, weights shared by the tile
        for (int r = tid >> 8; r < 10; r += RT_THREADS / 256) {
          const int col = tid & 255;
          const float* w = a.p.proj_w[r] + col;
          float acc[PB];
#pragma unroll
          for (int p = 0; p < PB; ++p) acc[p] = 0.f;
#pragma unroll 3
          for (int j = 0; j < 33; ++j) {
            const float wv = __ldg(w + (size_t)j * 256);
#pragma unroll
            for (int p = 0; p < PB; ++p) {
              RtPatient pt = rt_patient(pp, p, true);
              const float g = (j < 32) ? pt.dpose[r * 32 + j] * pt.rm[r] : pt.misc[32 + r];
              acc[p] = fmaf(g, wv, acc[p]);
            }
          }
#pragma unroll
          for (int p = 0; p < PB; ++p)
            if (p < np)
              a.d_route_embs[(size_t)r * a.d.emb_route_stride + (size_t)(b0 + p) * a.d.emb_batch_stride + col] = acc[p];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < RT_GSLOTS; ++j) {
    const int i = tid + RT_THREADS * j;
    if (i < K * 32 && a.dG) atomicAdd(a.dG + i, accG[j]);
  }
  if (tid < K && a.dbias) atomicAdd(a.dbias + tid, accB);
}

// dG[k][p] -> d pose_to_mc[m][p] += sum_k dG[k][p]*emb[k][m];  d emb[k][m] += sum_p dG[k][p]*Wmc[m][p]
// one thread per output element (MC*PC + K*MC of them)
__global__ void routing_head_grads_kernel(const float* dG, const float* pose_to_mc, const float* embedding, int K,
                                          float* d_pose_to_mc, float* d_embedding) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < MC * PC) {
    if (!d_pose_to_mc) return;
    const int m = i / PC, p = i % PC;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(dG[k * 32 + p], embedding[k * MC + m], acc);
    d_pose_to_mc[i] += acc;
  } else if (i < MC * PC + K * MC) {
    if (!d_embedding) return;
    const int j = i - MC * PC;
    const int k = j / MC, m = j % MC;
    float acc = 0.f;
    for (int p = 0; p < PC; ++p) acc = fmaf(dG[k * 32 + p], pose_to_mc[m * PC + p], acc);
    d_embedding[j] += acc;
  }
}

// fp16, reduction-index-contiguous copies of the two weight tensors the tensor-core paths read:
//   caps_wt[r][c][a] = w[r][a][c]            (10 x K*64 columns of 32)       one thread per column
//   caps_w16[r][a][c] = w[r][a][c]           (original layout, for the backward) one thread per 8 elements
//   proj_wb[r][n][k] = proj_w[r][n][k], n<33 (10 x 40 rows of 256, rest 0)   one thread per 8 elements
__global__ void routing_pack_kernel(mmr_routing_params p, int K, __half* caps_wt, __half* caps_w16, __half* proj_wb) {
  const int KD = K * 64;
  const long long n_caps = 10LL * KD;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_caps) {
    const int r = (int)(i / KD), c = (int)(i % KD);
    const float* w = p.caps_w + (size_t)r * 32 * KD + c;
    uint4* dst = reinterpret_cast<uint4*>(caps_wt + i * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 o;
      o.x = rt_pack2(w[(size_t)(8 * q + 0) * KD], w[(size_t)(8 * q + 1) * KD]);
      o.y = rt_pack2(w[(size_t)(8 * q + 2) * KD], w[(size_t)(8 * q + 3) * KD]);
      o.z = rt_pack2(w[(size_t)(8 * q + 4) * KD], w[(size_t)(8 * q + 5) * KD]);
      o.w = rt_pack2(w[(size_t)(8 * q + 6) * KD], w[(size_t)(8 * q + 7) * KD]);
      dst[q] = o;
    }
    return;
  }
  const long long n_w16 = 10LL * 32 * KD / 8;       // caps_w16[r][a][c] = w[r][a][c]: one thread per 8 elements
  if (i < n_caps + n_w16) {
    const long long e = (i - n_caps) * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(p.caps_w + e), x1 = *reinterpret_cast<const float4*>(p.caps_w + e + 4);
    uint4 o;
    o.x = rt_pack2(x0.x, x0.y); o.y = rt_pack2(x0.z, x0.w); o.z = rt_pack2(x1.x, x1.y); o.w = rt_pack2(x1.z, x1.w);
    reinterpret_cast<uint4*>(caps_w16)[i - n_caps] = o;
    return;
  }
  const long long j = i - n_caps - n_w16;
  if (proj_wb == nullptr || j >= 10LL * 40 * 32) return;
  const int row = (int)(j / 32), chunk = (int)(j % 32), r = row / 40, n = row % 40;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (n < 33) {
    const float* w = p.proj_w[r] + (size_t)n * 256 + chunk * 8;
    const float4 x0 = *reinterpret_cast<const float4*>(w), x1 = *reinterpret_cast<const float4*>(w + 4);
    o.x = rt_pack2(x0.x, x0.y); o.y = rt_pack2(x0.z, x0.w); o.z = rt_pack2(x1.x, x1.y); o.w = rt_pack2(x1.z, x1.w);
  }
  reinterpret_cast<uint4*>(proj_wb)[j] = o;
}

}  // namespace mmr
