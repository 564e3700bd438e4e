// Loss tail of the hot path: the step that consumes logits / alpha / R right after
// capsule routing, as device-side kernels that never synchronise the host (so the whole training step -- route
// fusion, routing, loss, backward, clip + AdamW + EMA -- is one CUDA graph):
//
//  * Mort   death logit + label smoothing + BCEWithLogits + route-entropy bonus + route-uniformity penalty
//           (MortModel/Paired_Cross_Attention/main.py:1753-1755, 3084-3126)
//  * Pheno  coerce_rc_to_report (main.py:1472-1564; two `.item()` host syncs in the reference), BCEWithLogits with
//           pos_weight, batch-mean routing entropy / uniformity terms (PhenoModel/.../main.py:2755-2812)
//
// One warp per patient, lane = label (K <= 32) or route: every per-patient reduction is lane-local or one shuffle
// tree, the [10 x K] routing slab of a patient is read once, coalesced.  Cross-patient reductions (BCE mean, the
// error maxima that select the coerce case, batch-mean routing coefficients) go through per-block partials in
// caller-owned scratch and are finished, in a fixed order, by the last block to arrive (ticket in the state struct),
// so results are bit-reproducible run to run.  All of it is a few hundred KB: launch-latency bound by construction.
#pragma once
#include "mmr_common.cuh"

namespace mmr {

constexpr int LOSS_WARPS = 8;            // patients in flight per block
constexpr int LOSS_PB = 32;              // patients per block
constexpr int LOSS_SLOTS = 336;          // doubles per block in scratch: [0] bce, [1] entropy, [2..12) alpha-dist column
                                         // sums, [12] / [13] max err over routes / labels, [14] max |sum_r R - 1|,
                                         // [15] rewritten logits, [16..336) sums of R over the block's patients
constexpr int LOSS_RK0 = 16;

struct LossArgs {
  int variant, B, K;
  const float* logits;                   // [B, K] (Mort: K = 2)
  const float* y;                        // Mort [B], Pheno [B, K]
  const float* pos_weight;               // [K] or null
  const float* prim_acts;                // [B, 10] or null
  const void* rc_raw; int rc_bf16;       // [B, 10, K] fp32 / bf16, or null
  const float* route_mask;               // [B, 10] or null
  float label_smoothing, lam_ent, lam_uni, atol;
  float* dlogits;                        // [B, K] or null
  float* rc_report;                      // [B, 10, K] or null
  mmr_loss_state* state;
  double* scratch;                       // [ceil(B / 32), LOSS_SLOTS]
};

__device__ __forceinline__ float loss_fix(float v, float posinf, float neginf, bool& was_finite) {
  was_finite = true;
  if (v != v) { was_finite = false; return 0.f; }
  if (isinf(v)) { was_finite = false; return v > 0.f ? posinf : neginf; }
  return v;
}
__device__ __forceinline__ float loss_ld_rc(const LossArgs& a, size_t i) {
  float v = a.rc_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(a.rc_raw)[i]) : reinterpret_cast<const float*>(a.rc_raw)[i];
  bool f;
  return loss_fix(v, 0.f, 0.f, f);       // torch.nan_to_num(rc_raw.detach().float(), 0, 0, 0)   main.py:1483
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// log(sigmoid(x)) = min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

// true for exactly one block per launch: the last one to have published its partials
__device__ __forceinline__ bool loss_last_block(mmr_loss_state* s) {
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&s->ticket, 1u);
    last = (t == gridDim.x - 1);
    if (last) s->ticket = 0;             // every block has arrived: safe to re-arm for the next launch
  }
  __syncthreads();
  if (last) __threadfence();
  return last != 0;
}

// Stage 1: BCE (+ gradient), Mort regularisers, Pheno coerce-case statistics.
__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_stage1_kernel(LossArgs a) {
  __shared__ double red[LOSS_WARPS][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool mort = a.variant == MMR_VARIANT_MORT;
  double bce = 0.0, ent = 0.0, pacol = 0.0;
  float err_r = 0.f, err_k = 0.f;
  int nonfinite = 0;
  const float inv_n = mort ? 1.0f / (float)a.B : 1.0f / ((float)a.B * (float)a.K);
  for (int b = blockIdx.x * LOSS_PB + warp; b < min(a.B, (int)(blockIdx.x + 1) * LOSS_PB); b += LOSS_WARPS) {
    if (mort) {
      // death_logit = logits[:, 1] - logits[:, 0]; y_s = y (1 - ls) + 0.5 ls; BCEWithLogits, mean over B
      if (lane == 0) {
        bool f0, f1;
        const float l0 = loss_fix(a.logits[2 * b], 1e4f, -1e4f, f0), l1 = loss_fix(a.logits[2 * b + 1], 1e4f, -1e4f, f1);
        nonfinite += (!f0) + (!f1);
        const float z = l1 - l0;
        float yv = a.y[b];
        if (a.label_smoothing > 0.f) yv = yv * (1.0f - a.label_smoothing) + 0.5f * a.label_smoothing;
        bce += (double)((1.0f - yv) * z - log_sigmoid(z));
        if (a.dlogits) {
          const float g = (sigmoidf(z) - yv) * inv_n;
          a.dlogits[2 * b] = f0 ? -g : 0.f;        // nan_to_num passes gradient only where its input was finite
          a.dlogits[2 * b + 1] = f1 ? g : 0.f;
        }
      }
      if (a.prim_acts) {
        // pa_dist = clamp_min(alpha, 1e-6) / clamp_min(sum_r, 1e-6)        main.py:3113-3114
        bool f;
        float pa = lane < NR ? fmaxf(loss_fix(a.prim_acts[(size_t)b * NR + lane], 1e4f, -1e4f, f), 1e-6f) : 0.f;
        const float s = fmaxf(warp_sum(pa), 1e-6f);
        const float pd = pa / s;
        if (lane < NR) {
          const float p = fmaxf(pd, 1e-12f);
          ent += (double)(-(p * logf(p)));
          pacol += (double)pd;
        }
      }
    } else {
      if (lane < a.K) {
        bool f;
        const float x = loss_fix(a.logits[(size_t)b * a.K + lane], 1e4f, -1e4f, f);
        nonfinite += !f;
        const float yv = a.y[(size_t)b * a.K + lane];
        const float lw = a.pos_weight ? (a.pos_weight[lane] - 1.0f) * yv + 1.0f : 1.0f;
        bce += (double)((1.0f - yv) * x - lw * log_sigmoid(x));
        if (a.dlogits) a.dlogits[(size_t)b * a.K + lane] = f ? ((1.0f - yv) - lw * (1.0f - sigmoidf(x))) * inv_n : 0.f;
      }
      if (a.rc_raw) {
        // s_over_routes [B,K], s_over_k [B,R] of the sanitised coefficients                       main.py:1488-1493
        float sr = 0.f;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const float v = lane < a.K ? loss_ld_rc(a, ((size_t)b * NR + r) * a.K + lane) : 0.f;
          sr += v;
          err_k = fmaxf(err_k, fabsf(warp_sum(v) - 1.0f));
        }
        if (lane < a.K) err_r = fmaxf(err_r, fabsf(sr - 1.0f));
      }
    }
  }
  // block partials; slot t: 0 bce, 1 entropy, 2..11 alpha-dist column sums, 12 / 13 error maxima, 15 rewritten logits
  bce = warp_sum_d(bce); ent = warp_sum_d(ent);
  err_r = warp_max(err_r); err_k = warp_max(err_k);
  const double nf = warp_sum_d((double)nonfinite);
  if (lane == 0) {
    red[warp][0] = bce; red[warp][1] = ent; red[warp][12] = (double)err_r; red[warp][13] = (double)err_k;
    red[warp][14] = 0.0; red[warp][15] = nf;
  }
  if (lane < NR) red[warp][2 + lane] = pacol;
  __syncthreads();
  double* mine = a.scratch + (size_t)blockIdx.x * LOSS_SLOTS;
  const bool is_max = threadIdx.x == 12 || threadIdx.x == 13;
  if (threadIdx.x < 16 && threadIdx.x != 14) {
    double v = red[0][threadIdx.x];
    for (int w = 1; w < LOSS_WARPS; ++w) v = is_max ? fmax(v, red[w][threadIdx.x]) : v + red[w][threadIdx.x];
    mine[threadIdx.x] = v;
  }
  if (!loss_last_block(a.state)) return;
  // ---- last block: finish the cross-patient reductions in block order ----
  __shared__ double tot[16];
  if (threadIdx.x < 16 && threadIdx.x != 14) {
    double v = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) {
      const double p = __ldcg(a.scratch + (size_t)k * LOSS_SLOTS + threadIdx.x);
      v = is_max ? fmax(v, p) : v + p;
    }
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mmr_loss_state* s = a.state;
    const float base = (float)(tot[0] * (double)inv_n);
    float e = 0.f, u = 0.f;
    if (mort && a.prim_acts) {
      if (a.lam_ent > 0.f) e = (float)(tot[1] / (double)a.B) * a.lam_ent;          // H = -(p log p).sum(1).mean()
      if (a.lam_uni > 0.f) {
        float acc = 0.f;
        for (int r = 0; r < NR; ++r) {
          const float d = (float)(tot[2 + r] / (double)a.B) - 1.0f / NR;                // (p_mean - 1/R)^2 summed over routes
          acc += d * d;
        }
        u = acc * a.lam_uni;
      }
    }
    s->base = base; s->ent = e; s->uni = u;
    s->loss = base - e + u;
    s->err_routes = (float)tot[12]; s->err_k = (float)tot[13];
    s->max_route_sum_err = 0.f;
    int info = 0;
    if (!mort && a.rc_raw) info = (float)tot[12] < a.atol ? 1 : ((float)tot[13] < a.atol ? 2 : 3);   // main.py:1495-1496
    s->info = info;
    s->nonfinite_logits = (int)tot[15];
  }
}

// Stage 2 (Pheno with routing coefficients): coerce_rc_to_report per patient, rc_report out, batch-mean entropy /
// uniformity terms.  info == 1: the coefficients already are p(route | phenotype) (what the capsule head returns);
// info == 3: forced normalisation over routes (clamp_min 0 first).  info == 2 (sums to one over labels): the
// reference's call `route_given_pheno(rc_raw_f, pa_f, route_mask=rm_f)` (main.py:1519) passes route_mask twice and
// raises TypeError; the kernel applies the forced normalisation and the host mirror raises when it reads `info`.
__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_stage2_kernel(LossArgs a) {
  __shared__ double red[LOSS_WARPS][NR][33];
  __shared__ float dev[LOSS_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int info = a.state->info;
  double acc[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) acc[r] = 0.0;
  float maxdev = 0.f;
  for (int b = blockIdx.x * LOSS_PB + warp; b < min(a.B, (int)(blockIdx.x + 1) * LOSS_PB); b += LOSS_WARPS) {
    float v[NR], m[NR];
    float denom = 0.f, avail = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      bool f;
      m[r] = a.route_mask ? loss_fix(a.route_mask[(size_t)b * NR + r], 0.f, 0.f, f) : 1.0f;
      v[r] = lane < a.K ? loss_ld_rc(a, ((size_t)b * NR + r) * a.K + lane) : 0.f;
      if (info != 1) v[r] = fmaxf(v[r], 0.f);
      if (a.route_mask) v[r] *= m[r];
      denom += v[r];
      avail += m[r];
    }
    const bool bad = !(denom == denom) || isinf(denom) || denom < 1e-8f;
    const float asum = fmaxf(avail, 1.0f);
    float d2 = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (bad) v[r] = a.route_mask ? m[r] / asum : 1.0f / NR;
      d2 += v[r];
    }
    d2 = fmaxf(d2, 1e-8f);
    float rsum = 0.f;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      v[r] = v[r] / d2;
      rsum += v[r];
      if (lane < a.K) {
        if (a.rc_report) a.rc_report[((size_t)b * NR + r) * a.K + lane] = v[r];
        acc[r] += (double)fminf(fmaxf(v[r], 1e-6f), 1.0f);                     // rc = routing_coef.clamp(1e-6, 1)   main.py:2795
      }
    }
    if (lane < a.K) maxdev = fmaxf(maxdev, fabsf(rsum - 1.0f));                // assert_routing_over_routes        main.py:261-276
  }
  maxdev = warp_max(maxdev);
  if (lane == 0) dev[warp] = maxdev;
#pragma unroll
  for (int r = 0; r < NR; ++r) red[warp][r][lane] = acc[r];
  __syncthreads();
  double* mine = a.scratch + (size_t)blockIdx.x * LOSS_SLOTS;
  for (int i = threadIdx.x; i < NR * 32; i += LOSS_WARPS * 32) {
    const int r = i >> 5, k = i & 31;
    double s = 0.0;
    for (int w = 0; w < LOSS_WARPS; ++w) s += red[w][r][k];
    mine[LOSS_RK0 + i] = s;
  }
  if (threadIdx.x == 0) {
    float dmax = dev[0];
    for (int w = 1; w < LOSS_WARPS; ++w) dmax = fmaxf(dmax, dev[w]);
    mine[14] = (double)dmax;
  }
  if (!loss_last_block(a.state)) return;
  // ---- last block: rc_bmean [R, K], H = mean_k(-sum_r m log m), U = mean_k(sum_r (m - 1/R)^2)      main.py:2798-2810
  __shared__ double hsum[LOSS_WARPS], usum[LOSS_WARPS], dsum[LOSS_WARPS];
  double h = 0.0, u = 0.0, dm = 0.0;
  for (int i = threadIdx.x; i < NR * 32; i += LOSS_WARPS * 32) {
    const int k = i & 31;
    if (k >= a.K) continue;
    double s = 0.0;
    for (unsigned blk = 0; blk < gridDim.x; ++blk) s += __ldcg(a.scratch + (size_t)blk * LOSS_SLOTS + LOSS_RK0 + i);
    const float mean = (float)(s / (double)a.B);
    h += (double)(-(mean * logf(mean)));
    const float d = mean - 1.0f / NR;
    u += (double)(d * d);
  }
  for (unsigned blk = threadIdx.x; blk < gridDim.x; blk += LOSS_WARPS * 32)
    dm = fmax(dm, __ldcg(a.scratch + (size_t)blk * LOSS_SLOTS + 14));
  h = warp_sum_d(h); u = warp_sum_d(u);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
  if (lane == 0) { hsum[warp] = h; usum[warp] = u; dsum[warp] = dm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double H = 0.0, U = 0.0, D = 0.0;
    for (int w = 0; w < LOSS_WARPS; ++w) { H += hsum[w]; U += usum[w]; D = fmax(D, dsum[w]); }
    mmr_loss_state* s = a.state;
    const float e = a.lam_ent > 0.f ? a.lam_ent * (float)(H / (double)a.K) : 0.f;
    const float un = a.lam_uni > 0.f ? a.lam_uni * (float)(U / (double)a.K) : 0.f;
    s->ent = e; s->uni = un;
    s->loss = s->base - e + un;
    s->max_route_sum_err = (float)D;
  }
}

// ---- evaluation-time routing statistics (first consumer) -------------------------------
// evaluate_epoch (MortModel/Paired_Cross_Attention/main.py:1916-1933) copies rc_raw, rc_report and prim_acts of every batch to
// the host and sums them there; here the three [10, K] sums (and the per-route activation sum) accumulate on the device and are
// read once per split.  Block = (route, label) column, threads stride over the patients.
struct RouteStatsArgs {
  const void* rc_raw; int rc_bf16;     // [B, 10, K]
  const float* rc_report;              // [B, 10, K] or null
  const float* prim_acts;              // [B, 10]
  int B, K;
  float* sums;                         // [3, 10, K] += (raw, report, raw * act)   then [10] += act
  unsigned long long* count;           // += B
};
__global__ void __launch_bounds__(256) route_stats_kernel(RouteStatsArgs a) {
  __shared__ float red[3][8];
  const int col = blockIdx.x, r = col / a.K;            // col = r * K + k; the last 10 blocks (col >= 10 K) sum the activations
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  if (col < 10 * a.K) {
    for (int b = threadIdx.x; b < a.B; b += 256) {
      const size_t i = (size_t)b * 10 * a.K + col;
      const float raw = a.rc_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(a.rc_raw)[i])
                                  : reinterpret_cast<const float*>(a.rc_raw)[i];
      s0 += raw;
      if (a.rc_report) s1 += a.rc_report[i];
      s2 += raw * a.prim_acts[(size_t)b * 10 + r];
    }
  } else {
    const int rr = col - 10 * a.K;
    for (int b = threadIdx.x; b < a.B; b += 256) s0 += a.prim_acts[(size_t)b * 10 + rr];
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    const int n = 10 * a.K;
    if (col < n) { a.sums[col] += t0; a.sums[n + col] += t1; a.sums[2 * n + col] += t2; }   // one block per address
    else a.sums[3 * n + (col - n)] += t0;
    if (col == 0 && a.count) *a.count += (unsigned long long)a.B;
  }
}

}  // namespace mmr
