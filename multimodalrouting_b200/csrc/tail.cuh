// The two steps either side of the route-fusion hot path, both pure
// HBM-bound element/row passes:
//
//  * producer epilogue   _clamp_norm + _safe_tensor + .float() of the encoder outputs
//                        (MortModel main.py:1772-1796; PhenoModel main.py:1445-1460 is nan_to_num only):
//                        one warp per token row, the row stays in registers between the norm and the scaling,
//                        one read + one write per element instead of the reference's ~8 elementwise passes;
//  * training tail       clip_grad_norm_ + finite check + AdamW + EMA (MortModel main.py:3143-3165, 58-108):
//                        multi-tensor kernels over (param, grad, exp_avg, exp_avg_sq, ema) quintuples -- one pass for
//                        the global gradient norm, one scalar kernel for clip / skip / bias corrections (everything
//                        stays on the device, so the step is CUDA-graph capturable and never syncs the host), one
//                        pass that reads 20 B and writes 16 B per parameter.
#pragma once
#include "mmr_common.cuh"

namespace mmr {

// ------------------------------------------------------------------------------ producer epilogue ---
constexpr int SAN_MAX_D = 1024;          // row width held in registers: 8 x float4 per lane
constexpr int SAN_WARPS = 8;

struct SanitizeArgs {
  const void* x; float* y;               // [rows, D] in (fp32 / bf16 / fp16), fp32 out
  const float* dy; float* dx;            // backward
  long long rows; int D;
  int mode;                              // 0: clamp-norm + nan_to_num(0, +-1e4)  1: nan_to_num(0, 0, 0) only
  float max_norm;
  unsigned long long* nonfinite;         // optional counter of non-finite entries seen AFTER the clamp (fwd)
};

template <class T> __device__ __forceinline__ float4 san_ld4(const T* p);
template <> __device__ __forceinline__ float4 san_ld4<float>(const float* p) { return Vec4<float>::ld(p); }
template <> __device__ __forceinline__ float4 san_ld4<bf16>(const bf16* p) { return Vec4<bf16>::ld(p); }
template <> __device__ __forceinline__ float4 san_ld4<__half>(const __half* p) { return Vec4<__half>::ld(p); }

__device__ __forceinline__ float san_fix(float v, float posinf, float neginf, int& bad) {
  if (v != v) { ++bad; return 0.f; }
  if (isinf(v)) { ++bad; return v > 0.f ? posinf : neginf; }
  return v;
}

// y = nan_to_num(x * min(1, max_norm / (||x||_2 + 1e-6)))      one warp per row
template <class T>
__global__ void __launch_bounds__(SAN_WARPS * 32) sanitize_fwd_kernel(SanitizeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * SAN_WARPS + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const T* x = reinterpret_cast<const T*>(a.x) + row * a.D;
  float* y = a.y + row * a.D;
  float4 v[SAN_MAX_D / 128];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < SAN_MAX_D / 128; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < a.D) {
      v[i] = san_ld4<T>(x + c);
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
  }
  float scale = 1.f;
  const bool mort = a.mode == 0;
  if (mort) {
    const float n = sqrtf(warp_sum(ss)) + 1e-6f;
    const float r = a.max_norm / n;
    scale = r > 1.0f ? 1.0f : r;          // torch.clamp(max=1): NaN propagates
  }
  const float pi = mort ? 1e4f : 0.f, ni = mort ? -1e4f : 0.f;
  int bad = 0;
#pragma unroll
  for (int i = 0; i < SAN_MAX_D / 128; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < a.D) {
      float4 o;
      if (mort) { o.x = v[i].x * scale; o.y = v[i].y * scale; o.z = v[i].z * scale; o.w = v[i].w * scale; }
      else o = v[i];
      o.x = san_fix(o.x, pi, ni, bad); o.y = san_fix(o.y, pi, ni, bad);
      o.z = san_fix(o.z, pi, ni, bad); o.w = san_fix(o.w, pi, ni, bad);
      Vec4<float>::st(y + c, o);
    }
  }
  if (a.nonfinite && bad) atomicAdd(a.nonfinite, (unsigned long long)bad);
}

// Backward of the forward above for rows of finite inputs (rows holding a NaN/Inf get dx = 0: the reference
// produces NaN gradients there and skips the optimizer step, main.py:3148-3151):
//   n = ||x|| + 1e-6;  n <  max_norm (scale 1, clamp inactive): dx = dy
//                      n >= max_norm: c = max_norm / n;  dx = c*dy - x * (max_norm / n^2) * (x . dy) / ||x||
template <class T>
__global__ void __launch_bounds__(SAN_WARPS * 32) sanitize_bwd_kernel(SanitizeArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * SAN_WARPS + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const T* x = reinterpret_cast<const T*>(a.x) + row * a.D;
  const float* dy = a.dy + row * a.D;
  float* dx = a.dx + row * a.D;
  float4 v[SAN_MAX_D / 128], g[SAN_MAX_D / 128];
  float ss = 0.f, dot = 0.f;
  bool finite = true;
#pragma unroll
  for (int i = 0; i < SAN_MAX_D / 128; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < a.D) {
      v[i] = san_ld4<T>(x + c);
      g[i] = Vec4<float>::ld(dy + c);
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
      dot += v[i].x * g[i].x + v[i].y * g[i].y + v[i].z * g[i].z + v[i].w * g[i].w;
    }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  finite = (ss == ss) && !isinf(ss);
  float c1 = 1.f, c2 = 0.f;
  if (a.mode == 0) {
    const float nrm = sqrtf(ss), n = nrm + 1e-6f;
    if (a.max_norm / n <= 1.0f) {         // clamp(max=1) passes the gradient where its input <= 1
      c1 = a.max_norm / n;
      c2 = -(a.max_norm / (n * n)) * dot / nrm;
    }
  }
  if (!finite) { c1 = 0.f; c2 = 0.f; }
#pragma unroll
  for (int i = 0; i < SAN_MAX_D / 128; ++i) {
    const int c = 4 * lane + 128 * i;
    if (c < a.D) {
      float4 o;
      if (finite) {
        o.x = fmaf(c2, v[i].x, c1 * g[i].x); o.y = fmaf(c2, v[i].y, c1 * g[i].y);
        o.z = fmaf(c2, v[i].z, c1 * g[i].z); o.w = fmaf(c2, v[i].w, c1 * g[i].w);
      } else {
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      Vec4<float>::st(dx + c, o);
    }
  }
}

// -------------------------------------------------------------------------- route mask from presence ---
// PhenoModel/Partial/Cross_Attention/routing_and_heads.py:10-64 (build_route_mask_from_presence) and main.py:109-132
// (build_route_mask_from_modalities): a route is allowed iff every modality it needs is present.
//   ROUTES = [L, N, I, LN, NL, LI, IL, NI, IN, LNI]; optional whole-batch dropped routes (MortModel main.py:3027-3033).
__global__ void route_mask_kernel(const float* hasL, const float* hasN, const float* hasI, int B, int drop_bits, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NR) return;
  const int b = i / NR, r = i % NR;
  // bit 0 = needs L, bit 1 = needs N, bit 2 = needs I
  const int needs[NR] = {1, 2, 4, 3, 3, 5, 5, 6, 6, 7};
  float m = 1.f;
  if ((needs[r] & 1) && hasL) m *= hasL[b];
  if ((needs[r] & 2) && hasN) m *= hasN[b];
  if ((needs[r] & 4) && hasI) m *= hasI[b];
  if ((drop_bits >> r) & 1) m = 0.f;
  out[i] = m;
}

// ---------------------------------------------------------------------------------- training tail ---
constexpr int OPT_NT = 384;              // tensors per launch (kernel-parameter table, 19.5 KB of the 32 KB sm_100 allows):
                                         // the 343 parameter tensors of the path go out in ONE launch
constexpr int OPT_CHUNK = 8192;          // elements per CTA
constexpr int OPT_THREADS = 256;

struct OptTable {
  float* p[OPT_NT]; const float* g[OPT_NT]; float* m[OPT_NT]; float* v[OPT_NT]; float* ema[OPT_NT];
  long long n[OPT_NT];
  int blk0[OPT_NT + 1];                  // first CTA of every tensor
  int nt;
};

using OptState = mmr_opt_state;         // device resident (include/mmr_b200.h)

// scalars are derived in double on the host and rounded to fp32 once, exactly where torch rounds its Python scalars
struct OptHyper { float beta2, eps, ema_decay, wd_factor, omb1, omb2, omd, step_size, bc2_sqrt; int has_ema; };

static inline OptHyper opt_hyper(const mmr_opt_hyper& h) {
  OptHyper o;
  o.beta2 = (float)h.beta2; o.eps = (float)h.eps; o.ema_decay = (float)h.ema_decay;
  o.wd_factor = (float)(1.0 - h.lr * h.weight_decay);
  o.omb1 = (float)(1.0 - h.beta1); o.omb2 = (float)(1.0 - h.beta2); o.omd = (float)(1.0 - h.ema_decay);
  o.step_size = 0.f; o.bc2_sqrt = 1.f; o.has_ema = h.has_ema;
  return o;
}

__device__ __forceinline__ int opt_find(const OptTable& t, int blk) {
  int lo = 0, hi = t.nt - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t.blk0[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(OPT_THREADS) opt_sqnorm_kernel(const __grid_constant__ OptTable t, OptState* s) {
  const int ti = opt_find(t, blockIdx.x);
  const long long e0 = (long long)(blockIdx.x - t.blk0[ti]) * OPT_CHUNK;
  const long long e1 = e0 + OPT_CHUNK < t.n[ti] ? e0 + OPT_CHUNK : t.n[ti];
  const float* g = t.g[ti];
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    long long i = e0 + 4 * threadIdx.x;
    for (; i + 3 < e1; i += 4 * OPT_THREADS) {
      const float4 x = *reinterpret_cast<const float4*>(g + i);
      acc += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    }
    for (; i < e1; ++i) acc += g[i] * g[i];       // at most one thread has a tail of < 4
  } else {
    for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) acc += g[i] * g[i];
  }
  acc = warp_sum(acc);
  __shared__ float red[OPT_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < OPT_THREADS / 32; ++w) tot += (double)red[w];
    atomicAdd(&s->sumsq, tot);
  }
}

// clip coefficient, skip flag, step count and bias corrections -- all on the device (no host sync)
__global__ void opt_prepare_kernel(OptState* s, double beta1, double beta2, float max_norm) {
  const double ss = s->sumsq;
  const float norm = (float)sqrt(ss);
  const bool bad = !(ss == ss) || isinf(ss);
  s->norm = norm;
  s->skip = bad ? 1 : 0;
  float clip = 1.f;
  if (max_norm > 0.f) {                           // torch.nn.utils.clip_grad_norm_: max_norm / (total_norm + 1e-6), <= 1
    clip = max_norm / (norm + 1e-6f);
    clip = clip > 1.f ? 1.f : clip;
  }
  s->clip = clip;
  if (!bad) {
    const int step = s->step + 1;
    s->step = step;
    s->bc1 = 1.0 - pow(beta1, (double)step);       // torch.optim.AdamW: Python-double bias corrections
    s->bc2_sqrt = sqrt(1.0 - pow(beta2, (double)step));
  }
  s->sumsq = 0.0;                                 // ready for the next step's norm pass
}

__device__ __forceinline__ void opt_update(float& p, float g, float& m, float& v, float* e, const OptState& s,
                                           const OptHyper& h) {
  g *= s.clip;                                    // clip_grad_norm_ scales the gradients in place
  p *= h.wd_factor;                               // AdamW decoupled weight decay: param.mul_(1 - lr * weight_decay)
  m = m + h.omb1 * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = v * h.beta2 + h.omb2 * g * g;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p + (-h.step_size) * (m / denom);           // param.addcdiv_(exp_avg, denom, value=-step_size)
  if (e) *e = *e * h.ema_decay + h.omd * p;       // EMA.update: sh.mul_(d).add_(v, alpha=1 - d)
}

__global__ void __launch_bounds__(OPT_THREADS) opt_step_kernel(const __grid_constant__ OptTable t,
                                                               const OptState* __restrict__ sp, OptHyper h, double lr) {
  const OptState s = *sp;
  if (s.skip) return;                             // non-finite gradients: the reference skips the step
  h.step_size = (float)(lr / s.bc1);              // step_size = lr / bias_correction1 (double, then fp32 scalar)
  h.bc2_sqrt = (float)s.bc2_sqrt;
  const int ti = opt_find(t, blockIdx.x);
  const long long e0 = (long long)(blockIdx.x - t.blk0[ti]) * OPT_CHUNK;
  const long long e1 = e0 + OPT_CHUNK < t.n[ti] ? e0 + OPT_CHUNK : t.n[ti];
  float* p = t.p[ti]; const float* g = t.g[ti]; float* m = t.m[ti]; float* v = t.v[ti];
  float* e = h.has_ema ? t.ema[ti] : nullptr;
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(e);
  long long i = e0 + threadIdx.x;
  if ((al & 15) == 0) {
    i = e0 + 4 * threadIdx.x;
    for (; i + 3 < e1; i += 4 * OPT_THREADS) {
      float4 P = *reinterpret_cast<float4*>(p + i), M = *reinterpret_cast<float4*>(m + i), V = *reinterpret_cast<float4*>(v + i);
      const float4 G = *reinterpret_cast<const float4*>(g + i);
      float4 E = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e) E = *reinterpret_cast<float4*>(e + i);
      opt_update(P.x, G.x, M.x, V.x, e ? &E.x : nullptr, s, h);
      opt_update(P.y, G.y, M.y, V.y, e ? &E.y : nullptr, s, h);
      opt_update(P.z, G.z, M.z, V.z, e ? &E.z : nullptr, s, h);
      opt_update(P.w, G.w, M.w, V.w, e ? &E.w : nullptr, s, h);
      *reinterpret_cast<float4*>(p + i) = P; *reinterpret_cast<float4*>(m + i) = M; *reinterpret_cast<float4*>(v + i) = V;
      if (e) *reinterpret_cast<float4*>(e + i) = E;
    }
    if (i >= e1) return;
    // the one thread that owns the < 4 element tail finishes it below
    for (; i < e1; ++i) opt_update(p[i], g[i], m[i], v[i], e ? e + i : nullptr, s, h);
    return;
  }
  for (; i < e1; i += OPT_THREADS) opt_update(p[i], g[i], m[i], v[i], e ? e + i : nullptr, s, h);
}

// EMA.update on its own (shadow = d * shadow + (1 - d) * value) for tensors the optimizer does not own
__global__ void __launch_bounds__(OPT_THREADS) ema_update_kernel(const __grid_constant__ OptTable t, float decay, float omd) {
  const int ti = opt_find(t, blockIdx.x);
  const long long e0 = (long long)(blockIdx.x - t.blk0[ti]) * OPT_CHUNK;
  const long long e1 = e0 + OPT_CHUNK < t.n[ti] ? e0 + OPT_CHUNK : t.n[ti];
  const float* p = t.p[ti]; float* e = t.ema[ti];
  for (long long i = e0 + threadIdx.x; i < e1; i += OPT_THREADS) e[i] = e[i] * decay + omd * p[i];
}

}  // namespace mmr
