// Row-wise (warp-per-token) kernels of the route-fusion path: embedding + positional add,
// LayerNorm forward/backward with fused masking / residual-gradient accumulation, masked-mean
// pooling, column sums for bias gradients and weight packing / gradient unfolding.
#pragma once
#include "mmr_common.cuh"

namespace mmr {

constexpr int ROWS_PER_BLOCK = 8;   // 8 warps, one token row each

// route index (ROUTES order L,N,I,LN,NL,LI,IL,NI,IN,LNI) of direction d = LN,LI,NL,NI,IL,IN
__host__ __device__ inline int route_of_dir(int d) {
  const int t[6] = {3, 5, 4, 7, 6, 8};
  return t[d];
}
// pair (0: LN/NL, 1: LI/IL, 2: NI/IN) and half (0 first operand of the concat) of direction d
__host__ __device__ inline int pair_of_dir(int d) { const int t[6] = {0, 1, 0, 2, 1, 2}; return t[d]; }
__host__ __device__ inline int half_of_dir(int d) { const int t[6] = {0, 0, 1, 0, 1, 1}; return t[d]; }

// -------------------------------------------------------------------------------------------
// Row plan of the packed query space (see Segs in mmr_common.cuh): counts the valid tokens of every patient (mask != 0), scans
// the counts into patient offsets and fills the token <-> row maps.  pack = 0 writes the dense identity plan
// (row = b * T + t for every token).
struct RowPlanArgs {
  const float* mask[NMOD];   // [B, T] or null
  int T[NMOD];
  int B, pack;
  int* nv;                   // [8]: nv[d] = valid rows of direction d (query modality d >> 1)
  int* poff[NMOD]; int* tokrow[NMOD]; int* rowpat[NMOD];
};

// Three launches (round 2): counting and the map fill get one warp PER PATIENT across the grid, only the scan of the B counts
// is one block per modality.  The single-block form walked 16 patients per warp twice, each behind a dependent load of its
// mask row: 30 us at B = 512 at the head of every forward call (ncu: 3 CTAs, 8 long-scoreboard stall cycles per instruction).
// grid (ceil(B / 32), NMOD), 1024 threads
__global__ void __launch_bounds__(1024) rowplan_count_kernel(RowPlanArgs a) {
  const int m = blockIdx.y, T = a.T[m], lane = threadIdx.x & 31;
  const int b = blockIdx.x * 32 + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const float* mk = a.mask[m];
  int c = T;
  if (a.pack && mk != nullptr) {
    c = 0;
    for (int t = lane; t < T; t += 32) c += mk[(size_t)b * T + t] != 0.f ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) a.poff[m][b + 1] = c;
}

// grid NMOD, 1024 threads: inclusive scan of poff[1..B] in chunks of 1024, nv
__global__ void __launch_bounds__(1024) rowplan_scan_kernel(RowPlanArgs a) {
  __shared__ int part[1024];
  __shared__ int base_s;
  const int m = blockIdx.x, B = a.B, tid = threadIdx.x;
  int* poff = a.poff[m];
  if (tid == 0) { poff[0] = 0; base_s = 0; }
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += 1024) {
    const int b = b0 + tid;
    int v = b < B ? poff[b + 1] : 0;
    part[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int add = tid >= o ? part[tid - o] : 0;
      __syncthreads();
      part[tid] += add;
      __syncthreads();
    }
    if (b < B) poff[b + 1] = base_s + part[tid];
    __syncthreads();
    if (tid == 1023) base_s += part[1023];
    __syncthreads();
  }
  if (tid == 0) { a.nv[2 * m] = base_s; a.nv[2 * m + 1] = base_s; if (m == 0) { a.nv[6] = 0; a.nv[7] = 0; } }
}

// grid (ceil(B / 32), NMOD), 1024 threads: token <-> row maps (a warp per patient, tokens in order)
__global__ void __launch_bounds__(1024) rowplan_fill_kernel(RowPlanArgs a) {
  const int m = blockIdx.y, T = a.T[m], lane = threadIdx.x & 31;
  const int b = blockIdx.x * 32 + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const float* mk = a.mask[m];
  const bool dense = !a.pack || mk == nullptr;
  int row = a.poff[m][b];
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool ok = t < T && (dense || mk[(size_t)b * T + t] != 0.f);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int rank = __popc(bal & ((1u << lane) - 1u));
    if (t < T) a.tokrow[m][(size_t)b * T + t] = ok ? row + rank : -1;
    if (ok) a.rowpat[m][row + rank] = b;
    row += __popc(bal);
  }
}

// -------------------------------------------------------------------------------------------
// embed: E = 16*p + pos[t]  (transformer.py:63-72);  unimodal encoder (0 layers) output
// U = LN_uni(E*m)*m (transformer.py:77-79,108-113); normalised key/value stream XH = (E-mean)*rstd
// (the affine part of each layer's LN0 is folded into the K/V weights); and for the two
// directions that use this modality as the query: x0 = E*m, h0 = LN0_{layer0}(x0)*m.
struct EmbedArgs {
  Segs mod, q;
  const float* src[NMOD];      // [B*T, 256] fp32 (inputs or projected inputs)
  const float* mask[NMOD];     // [B*T] or null
  const float* pos;            // [maxT, 256]
  const float* uni_g[NMOD]; const float* uni_b[NMOD];
  const float* ln0_g[NDIR]; const float* ln0_b[NDIR];   // layer-0 LN0 of each direction
  void* xh; float* rstd_e; float* u;
  float* xin0; void* h0; float* stat0; float* maskq;
  const int* tokrow[NMOD];     // token -> row inside the query segment (-1: padded token, no query row)
};

template <class CT>
__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= a.mod.row0[a.mod.n]) return;
  const int m = seg_of_row(a.mod, r);
  const int local = r - a.mod.row0[m];
  const bool valid = local < a.mod.rows[m];
  Row8 e, z;
#pragma unroll
  for (int i = 0; i < 8; ++i) { e.v[i] = 0.f; z.v[i] = 0.f; }
  float mval = 0.f, mean = 0.f, rstd = 0.f;
  Row8 xh = z, u = z;
  if (valid) {
    const int t = local % a.mod.T[m];
    Row8 x = row_load<float>(a.src[m] + (size_t)local * D, lane);
    Row8 p = row_load<float>(a.pos + (size_t)t * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) e.v[i] = EMBED_SCALE * x.v[i] + p.v[i];
    mval = a.mask[m] ? a.mask[m][local] : 1.f;
    row_stats(e, mean, rstd);
#pragma unroll
    for (int i = 0; i < 8; ++i) xh.v[i] = (e.v[i] - mean) * rstd;
    // unimodal route: LN(E*m)*m
    Row8 em;
#pragma unroll
    for (int i = 0; i < 8; ++i) em.v[i] = e.v[i] * mval;
    float mu, ru;
    row_stats(em, mu, ru);
    Row8 g = row_load<float>(a.uni_g[m], lane), bb = row_load<float>(a.uni_b[m], lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) u.v[i] = ((em.v[i] - mu) * ru * g.v[i] + bb.v[i]) * mval;
  }
  row_store<CT>(reinterpret_cast<CT*>(a.xh) + (size_t)r * D, lane, xh);
  row_store<float>(a.u + (size_t)r * D, lane, u);
  if (lane == 0) a.rstd_e[r] = rstd;
  // query streams of directions 2m and 2m+1
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int d = 2 * m + k;
    const int qrl = valid ? a.tokrow[m][local] : -1;
    if (qrl < 0) continue;      // padded token: no query row (the rows [nv, pad) of the segment are zeroed by zero_pad)
    const int qr = a.q.row0[d] + qrl;
    Row8 x0 = z, h0 = z;
    float mu = 0.f, ru = 0.f;
    {
#pragma unroll
      for (int i = 0; i < 8; ++i) x0.v[i] = e.v[i] * mval;
      row_stats(x0, mu, ru);
      Row8 g = row_load<float>(a.ln0_g[d], lane), bb = row_load<float>(a.ln0_b[d], lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) h0.v[i] = ((x0.v[i] - mu) * ru * g.v[i] + bb.v[i]) * mval;
    }
    row_store<float>(a.xin0 + (size_t)qr * D, lane, x0);
    row_store<CT>(reinterpret_cast<CT*>(a.h0) + (size_t)qr * D, lane, h0);
    if (lane == 0) {
      a.stat0[2 * (size_t)qr] = mu;
      a.stat0[2 * (size_t)qr + 1] = ru;
      a.maskq[qr] = mval;
    }
  }
}

// -------------------------------------------------------------------------------------------
// LayerNorm forward over the query row space: out = (LN(x)*gamma+beta)*maskq; stats saved.
// With `delta` (the bias-added output of out_proj / fc2 in the compute type) the residual add and
// query masking of transformer.py:196-199,211-215 are fused in: x = (x_prev + delta)*maskq is
// written to x_out before being normalised.
struct LnFwdArgs {
  Segs q;
  const float* x; const float* maskq;
  const void* delta; float* x_out;
  const float* gamma[NDIR]; const float* beta[NDIR];
  void* out; float* stat;
};

template <class OT, class CT>
__global__ void __launch_bounds__(256) ln_rows_fwd_kernel(LnFwdArgs a) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= a.q.row0[a.q.n]) return;
  const int d = seg_of_row(a.q, r);
  const int local = r - a.q.row0[d];
  if (local >= seg_rows(a.q, d)) {       // not a query row: zero inside the tile-padding zone, untouched beyond it
    if (local < seg_rows_z(a.q, d)) {
      Row8 zr;
#pragma unroll
      for (int i = 0; i < 8; ++i) zr.v[i] = 0.f;
      if (a.delta) row_store<float>(a.x_out + (size_t)r * D, lane, zr);
      row_store<OT>(reinterpret_cast<OT*>(a.out) + (size_t)r * D, lane, zr);
      if (lane == 0) { a.stat[2 * (size_t)r] = 0.f; a.stat[2 * (size_t)r + 1] = 0.f; }
    }
    return;
  }
  const float mval = a.maskq[r];
  Row8 x = row_load<float>(a.x + (size_t)r * D, lane);
  if (a.delta) {
    Row8 dl = row_load<CT>(reinterpret_cast<const CT*>(a.delta) + (size_t)r * D, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) x.v[i] = (x.v[i] + dl.v[i]) * mval;
    row_store<float>(a.x_out + (size_t)r * D, lane, x);
  }
  float mean, rstd;
  row_stats(x, mean, rstd);
  Row8 g = row_load<float>(a.gamma[d], lane), b = row_load<float>(a.beta[d], lane), o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = ((x.v[i] - mean) * rstd * g.v[i] + b.v[i]) * mval;
  row_store<OT>(reinterpret_cast<OT*>(a.out) + (size_t)r * D, lane, o);
  if (lane == 0) { a.stat[2 * (size_t)r] = mean; a.stat[2 * (size_t)r + 1] = rstd; }
}

// -------------------------------------------------------------------------------------------
// LayerNorm backward over the query row space.
//   dh      : gradient wrt (LN(x)*gamma+beta)*maskq, already multiplied by maskq
//             POOLED: dh = dz[dir][b]*maskq^2/cnt[b] is synthesised from the pooled gradient
//   g_out   = (g_in + dLN/dx) * maskq   (fp32) and a compute-type copy gc_out
//   dgamma[dir] += sum dh*xhat, dbeta[dir] += sum dh, dbias[dir] += sum g_out (all optional)
struct LnBwdArgs {
  Segs q;
  const void* dh;                 // CT [MQ,256] (non-pooled)
  const float* dz[NDIR];          // pooled: fp32 [B,256] per direction (effective route gradient)
  const float* dz2[NDIR];         // pooled: second addend [B, ld2] (pair-projection gradient) or null
  int ld2;
  const float* cnt[NDIR];         // pooled: fp32 [B]
  const float* x; const float* stat; const float* maskq;
  const float* gamma[NDIR];
  const float* g_in;              // fp32 [MQ,256] or null
  float* g_out; void* gc_out;
  float* dgamma[NDIR]; float* dbeta[NDIR]; float* dbias[NDIR];
};

#ifndef LN_BWD_MINB
#define LN_BWD_MINB 3
#endif
template <class CT, bool POOLED>
__global__ void __launch_bounds__(256, LN_BWD_MINB) ln_rows_bwd_kernel(LnBwdArgs a) {
  constexpr int RPB = 64;   // rows per block (never straddles a 128-aligned segment)
  __shared__ float red[8][3][256];
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * RPB;
  const int d = seg_of_row(a.q, r0);
  float ag[8], ab[8], as[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; as[i] = 0.f; }
  Row8 gm = row_load<float>(a.gamma[d], lane);
  for (int rr = warp; rr < RPB; rr += 8) {
    const int r = r0 + rr;
    const int local = r - a.q.row0[d];
    const bool valid = local < seg_rows(a.q, d);
    if (!valid && local >= seg_rows_z(a.q, d)) continue;     // beyond the tile-padding zone: never read by anyone
    Row8 out;
#pragma unroll
    for (int i = 0; i < 8; ++i) out.v[i] = 0.f;
    if (valid) {
      // Every operand of the row is requested BEFORE the first use.  The kernel used to load the keep-mask, branch on it,
      // load dh / x / statistics, reduce, and only then load the incoming residual gradient: three dependent memory round
      // trips per row (ncu: 61 % of the stall samples on those two waits, 4.5 TB/s).  With packed query rows every row is a
      // valid token, so the speculative loads are never wasted; in the dense layout a padded token costs its loads.
      const float mval = a.maskq[r];
      Row8 x = row_load<float>(a.x + (size_t)r * D, lane);
      const float mean = a.stat[2 * (size_t)r], rstd = a.stat[2 * (size_t)r + 1];
      Row8 gi;
#pragma unroll
      for (int i = 0; i < 8; ++i) gi.v[i] = 0.f;
      if (a.g_in) gi = row_load<float>(a.g_in + (size_t)r * D, lane);
      Row8 dh;
      if (POOLED) {
        const int b = seg_row_patient(a.q, d, local);
        dh = row_load<float>(a.dz[d] + (size_t)b * D, lane);
        if (a.dz2[d]) {
          Row8 e2 = row_load<float>(a.dz2[d] + (size_t)b * a.ld2, lane);
#pragma unroll
          for (int i = 0; i < 8; ++i) dh.v[i] += e2.v[i];
        }
        const float sc = mval * mval / a.cnt[d][b];
#pragma unroll
        for (int i = 0; i < 8; ++i) dh.v[i] *= sc;
      } else {
        dh = row_load<CT>(reinterpret_cast<const CT*>(a.dh) + (size_t)r * D, lane);
      }
      if (mval != 0.f) {
        float s1 = 0.f, s2 = 0.f;
        float xh[8], gx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[i] = (x.v[i] - mean) * rstd;
          gx[i] = dh.v[i] * gm.v[i];
          s1 += gx[i];
          s2 += gx[i] * xh[i];
          ag[i] += dh.v[i] * xh[i];
          ab[i] += dh.v[i];
        }
        s1 = warp_sum(s1) * (1.0f / D);
        s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          out.v[i] = (rstd * (gx[i] - s1 - xh[i] * s2) + gi.v[i]) * mval;
          as[i] += out.v[i];
        }
      }
    }
    row_store<float>(a.g_out + (size_t)r * D, lane, out);
    row_store<CT>(reinterpret_cast<CT*>(a.gc_out) + (size_t)r * D, lane, out);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = row_col(lane, i);
    red[warp][0][c] = ag[i]; red[warp][1][c] = ab[i]; red[warp][2][c] = as[i];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { t0 += red[w][0][c]; t1 += red[w][1][c]; t2 += red[w][2][c]; }
  if (a.dgamma[d]) atomicAdd(a.dgamma[d] + c, t0);
  if (a.dbeta[d]) atomicAdd(a.dbeta[d] + c, t1);
  if (a.dbias[d]) atomicAdd(a.dbias[d] + c, t2);
}

// -------------------------------------------------------------------------------------------
// Masked-mean pooling (mult_model.py:84-90) for the 3 unimodal + 6 cross-modal routes.
// Writes routes[route][b][:] and the concatenated pair operand zcat[pair][b][half*256 + :].
struct PoolArgs {
  Segs mod, q;
  const float* u; const float* y;         // unimodal outputs (modality rows), final-LN outputs (q rows)
  const float* mask[NMOD];
  const float* maskq;                     // [MQ] keep-mask value of every query row
  float* routes;                          // [10,B,256]
  float* zcat;                            // [3,B,512]
  float* cnt;                             // [3,B]
  int B;
};

__global__ void __launch_bounds__(256) pool_fwd_kernel(PoolArgs a) {
  const int b = blockIdx.x, which = blockIdx.y, c = threadIdx.x;   // which: 0..2 unimodal, 3..8 directions
  int mod, T, route;
  const float* src;
  const float* mk;
  if (which < 3) {
    mod = which; T = a.mod.T[mod]; route = which;
    src = a.u + ((size_t)a.mod.row0[mod] + (size_t)b * T) * D;
    mk = a.mask[mod] ? a.mask[mod] + (size_t)b * T : nullptr;
  } else {                                  // the patient's (packed) query rows; masked tokens have no row and weight 0
    const int d = which - 3;
    int start;
    seg_patient(a.q, d, b, start, T);
    mod = dir_qmod(d); route = route_of_dir(d);
    src = a.y + ((size_t)a.q.row0[d] + start) * D;
    mk = a.maskq + a.q.row0[d] + start;
  }
  float acc = 0.f, cnt = 0.f;
  for (int t0 = 0; t0 < T; t0 += 8) {      // eight tokens' loads in flight per thread; same summation order as a plain loop
    float v[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool in = t0 + j < T;
      m[j] = in ? (mk ? mk[t0 + j] : 1.f) : 0.f;
      v[j] = in ? src[(size_t)(t0 + j) * D + c] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc += v[j] * m[j]; cnt += m[j]; }
  }
  cnt = fmaxf(cnt, 1.0f);
  const float z = acc / cnt;
  a.routes[((size_t)route * a.B + b) * D + c] = z;
  if (which < 3) {
    if (c == 0) a.cnt[mod * a.B + b] = cnt;
  } else {
    const int d = which - 3;
    a.zcat[((size_t)pair_of_dir(d) * a.B + b) * 512 + half_of_dir(d) * D + c] = z;
  }
}

// -------------------------------------------------------------------------------------------
// Embedding backward (modality row space): collects the gradients of E from the two query
// streams, the K/V stream (through the affine-free normalisation) and the unimodal route.
struct EmbedBwdArgs {
  Segs mod, q, kv;
  const void* xh; const float* rstd_e;
  const float* mask[NMOD];
  const float* g0;                 // fp32 [MQ,256] gradient wrt x0 of every direction (masked)
  const float* dxh;                // fp32 [MK,256] gradient wrt XH per direction (kv rows)
  const float* dz_uni[NMOD];       // fp32 [B,256] gradient of the unimodal route embeddings
  const float* cnt;                // [3,B]
  const float* uni_g[NMOD];
  float* d_uni_g[NMOD]; float* d_uni_b[NMOD];
  float* dsrc[NMOD];               // fp32 [B*T,256] gradient wrt the (projected) inputs, or null
  const int* tokrow[NMOD];         // token -> row inside the query segment (-1: padded token)
  int B;
};

template <class CT>
__global__ void __launch_bounds__(256, 3) embed_bwd_kernel(EmbedBwdArgs a) {
  constexpr int RPB = 64;
  __shared__ float red[8][2][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r0 = blockIdx.x * RPB;
  const int m = seg_of_row(a.mod, r0);
  const int T = a.mod.T[m];
  // directions whose key/value modality is m
  int kd[2]; { int n = 0; for (int d = 0; d < NDIR; ++d) if (dir_kmod(d) == m) kd[n++] = d; }
  float ag[8], ab[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ag[i] = 0.f; ab[i] = 0.f; }
  Row8 gm = row_load<float>(a.uni_g[m], lane);
  for (int rr = warp; rr < RPB; rr += 8) {
    const int r = r0 + rr;
    const int local = r - a.mod.row0[m];
    if (local >= a.mod.rows[m]) continue;
    const int b = local / T;
    // every load that does not depend on another load of the row is requested here, before the first use (the keep-mask,
    // the token -> query-row map and the pooled-route gradient used to head three separate dependent round trips)
    const float mval = a.mask[m] ? a.mask[m][local] : 1.f;
    const int qrl = a.tokrow[m][local];
    Row8 xh = row_load<CT>(reinterpret_cast<const CT*>(a.xh) + (size_t)r * D, lane);
    const float rstd = a.rstd_e[r];
    Row8 k0 = row_load<float>(a.dxh + ((size_t)a.kv.row0[kd[0]] + local) * D, lane);
    Row8 k1 = row_load<float>(a.dxh + ((size_t)a.kv.row0[kd[1]] + local) * D, lane);
    Row8 dz;
    float cntb = 1.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dz.v[i] = 0.f;
    if (a.dz_uni[m]) {
      dz = row_load<float>(a.dz_uni[m] + (size_t)b * D, lane);
      cntb = a.cnt[m * a.B + b];
    }
    Row8 qa, qb;
    if (qrl >= 0) {
      qa = row_load<float>(a.g0 + ((size_t)a.q.row0[2 * m] + qrl) * D, lane);
      qb = row_load<float>(a.g0 + ((size_t)a.q.row0[2 * m + 1] + qrl) * D, lane);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { qa.v[i] = 0.f; qb.v[i] = 0.f; }
    }
    float gt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) gt[i] = k0.v[i] + k1.v[i];
    if (mval != 0.f && a.dz_uni[m]) {
      const float sc = mval * mval / cntb;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dy = dz.v[i] * sc;
        ag[i] += dy * xh.v[i];
        ab[i] += dy;
        gt[i] += dy * gm.v[i] * mval;
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1 += gt[i]; s2 += gt[i] * xh.v[i]; }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
    Row8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o.v[i] = EMBED_SCALE * ((qa.v[i] + qb.v[i]) * mval + rstd * (gt[i] - s1 - xh.v[i] * s2));
    if (a.dsrc[m]) row_store<float>(a.dsrc[m] + (size_t)local * D, lane, o);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = row_col(lane, i);
    red[warp][0][c] = ag[i]; red[warp][1][c] = ab[i];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float t0 = 0.f, t1 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { t0 += red[w][0][c]; t1 += red[w][1][c]; }
  if (a.d_uni_g[m]) atomicAdd(a.d_uni_g[m] + c, t0);
  if (a.d_uni_b[m]) atomicAdd(a.d_uni_b[m] + c, t1);
}

// -------------------------------------------------------------------------------------------
// Column sums over row segments: out[seg][c] += sum_r src[r, col0 + c].  blockDim = 256 columns,
// grid = (ceil(ncols/256), padded_rows/128).
struct ColsumArgs {
  Segs segs;
  const void* src; int ld; int col0; int ncols;
  float* out[6];
  float scale;
  int colblock;     // > 0: single segment, column c accumulates into out[c / colblock][c % colblock]
};
template <class T>
__global__ void __launch_bounds__(256) colsum_kernel(ColsumArgs a) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int r0 = blockIdx.y * 128;
  if (r0 >= a.segs.row0[a.segs.n]) return;
  const int seg = seg_of_row(a.segs, r0);
  if (c >= a.ncols) return;
  float* dst = a.colblock > 0 ? a.out[c / a.colblock] : a.out[seg];
  if (dst == nullptr) return;
  dst += a.colblock > 0 ? c % a.colblock : c;
  const int r_end = min(r0 + 128, a.segs.row0[seg] + seg_rows(a.segs, seg));
  const T* p = reinterpret_cast<const T*>(a.src) + a.col0 + c;
  float s = 0.f;
  for (int r = r0; r < r_end; ++r) s += to_f<T>(p[(size_t)r * a.ld]);
  atomicAdd(dst, s * a.scale);
}

// Zero the tile-padding rows [seg_rows, seg_rows_z) of every segment of a [*, ld_bytes] buffer (grid.x rows per pass).
__global__ void zero_pad_rows_kernel(Segs s, uint8_t* buf, size_t ld_bytes) {
  const int seg = blockIdx.y;
  const int r_begin = seg_rows(s, seg), r_end = seg_rows_z(s, seg);
  for (int pr = r_begin + blockIdx.x; pr < r_end; pr += gridDim.x) {
    uint4* row = reinterpret_cast<uint4*>(buf + ((size_t)s.row0[seg] + pr) * ld_bytes);
    for (size_t i = threadIdx.x; i < ld_bytes / 16; i += blockDim.x) row[i] = make_uint4(0, 0, 0, 0);
  }
}

// -------------------------------------------------------------------------------------------
// Weight packing: dst[r][c] = scale * src[r][c] * colscale[c]  (compute type), dstT = transpose.
struct PackJob {
  const float* src; int rows, cols, src_ld;
  const float* colscale; float scale;
  void* dst; int dst_ld;
  void* dstT; int dstT_ld;
};
struct PackJobs { PackJob j[60]; int n; };

template <class CT>
__global__ void __launch_bounds__(256) pack_kernel(PackJobs jobs) {
  __shared__ float tile[32][33];
  const PackJob& j = jobs.j[blockIdx.y];
  const int tiles_c = j.cols / 32;                 // grid.x = 256 tiles covers every job ([1024,256] / [256,1024] max)
  const int r0 = ((int)blockIdx.x / tiles_c) * 32, c0 = ((int)blockIdx.x % tiles_c) * 32;
  if (r0 >= j.rows) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  CT* dst = reinterpret_cast<CT*>(j.dst);
  CT* dstT = reinterpret_cast<CT*>(j.dstT);
  for (int rr = ty; rr < 32; rr += 8) {
    float v = j.src[(size_t)(r0 + rr) * j.src_ld + c0 + tx] * j.scale;
    if (j.colscale) v *= j.colscale[c0 + tx];
    tile[rr][tx] = v;
    if (dst) dst[(size_t)(r0 + rr) * j.dst_ld + c0 + tx] = from_f<CT>(v);
  }
  __syncthreads();
  if (dstT)
    for (int cc = ty; cc < 32; cc += 8) dstT[(size_t)(c0 + cc) * j.dstT_ld + r0 + tx] = from_f<CT>(tile[tx][cc]);
}

// bias fold: dst[o] = scale * (b[o] + sum_i W[o][i] * beta[i]); one warp per output row.
struct BiasJob { const float* w; int ld; const float* b; const float* beta; float scale; float* dst; int rows; };
struct BiasJobs { BiasJob j[60]; int n; };
__global__ void __launch_bounds__(256) bias_fold_kernel(BiasJobs jobs) {
  const BiasJob& j = jobs.j[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= j.rows) return;
  float s = 0.f;
  if (j.beta) {
    Row8 w = row_load<float>(j.w + (size_t)o * j.ld, lane), be = row_load<float>(j.beta, lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w.v[i] * be.v[i];
    s = warp_sum(s);
  }
  if (lane == 0) j.dst[o] = j.scale * (j.b[o] + s);
}

// Gradient unfolding for one (direction, layer): packed-weight gradients -> in_proj_{weight,bias},
// plus the K/V-path contribution to layer_norms.0.{weight,bias}.
//   Wq' = s*Wq                      -> dWq = s*dWq',  dbq = s*dbq'
//   Wkv' = Wkv*diag(gamma0)         -> dWkv = dWkv'*diag(gamma0), dgamma0 += colsum(dWkv' .* Wkv)
//   bkv' = bkv + Wkv*beta0          -> dbkv = dbkv', dbeta0 += Wkv^T dbkv', dWkv += dbkv' beta0^T
struct UnfoldJob {
  const float* dwq; const float* dbq;        // [256,256], [256]
  const float* dwkv; const float* dbkv;      // [512,256], [512]
  const float* w_in; const float* gamma0; const float* beta0;   // in_proj_weight [768,256], ln0 weight/bias [256]
  float* g_w_in; float* g_b_in; float* g_gamma0; float* g_beta0;
};
struct UnfoldJobs { UnfoldJob j[24]; int n; float scaling; };
constexpr int UNFOLD_ROWS = 32;   // in_proj rows per block; grid = (768 / 32, jobs)
__global__ void __launch_bounds__(256) unfold_kernel(UnfoldJobs jobs) {
  const UnfoldJob& j = jobs.j[blockIdx.y];
  const int c = threadIdx.x;
  const float s = jobs.scaling;
  const int r0 = blockIdx.x * UNFOLD_ROWS;        // row of in_proj_weight [768,256]
  if (r0 < D) {                                   // query block: Wq' = s*Wq
    if (j.g_w_in)
#pragma unroll 4
      for (int o = r0; o < r0 + UNFOLD_ROWS; ++o) j.g_w_in[(size_t)o * D + c] += s * j.dwq[(size_t)o * D + c];
  } else {                                        // key / value blocks (LN0 affine folded into the weights)
    const float gm = j.gamma0[c], be = j.beta0[c];
    float dg = 0.f, db = 0.f;
#pragma unroll 4
    for (int o = r0 - D; o < r0 - D + UNFOLD_ROWS; ++o) {
      const float dw = j.dwkv[(size_t)o * D + c];
      const float w = j.w_in[(size_t)(D + o) * D + c];
      const float dbo = j.dbkv[o];
      if (j.g_w_in) j.g_w_in[(size_t)(D + o) * D + c] += dw * gm + dbo * be;
      dg += dw * w;
      db += dbo * w;
    }
    if (j.g_gamma0) atomicAdd(j.g_gamma0 + c, dg);
    if (j.g_beta0) atomicAdd(j.g_beta0 + c, db);
  }
  if (blockIdx.x == 0 && j.g_b_in) {
    j.g_b_in[c] += s * j.dbq[c];
    j.g_b_in[D + c] += j.dbkv[c];
    j.g_b_in[2 * D + c] += j.dbkv[D + c];
  }
}

// out[i] = a[i] + b[i] (b may be null), n elements; used to form effective route gradients.
__global__ void add_kernel(float* out, const float* a, const float* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + (b ? b[i] : 0.f);
}

}  // namespace mmr
