// Per-patient multi-head cross attention (multihead_attention.py:93-143), all six directions in
// one launch.  Sequences are short (16-49 tokens at MIMIC shapes, <=512 in the stress config) so one
// thread owns one (head, query) row: K/V chunks are staged in shared memory as fp32 and read by
// broadcast; softmax is online in fp32.  Key padding uses masked_fill(finfo(dtype).min) semantics:
// a patient whose keys are all padded attends uniformly over the padded rows.
#pragma once
#include <float.h>

#include "mmr_common.cuh"

namespace mmr {

constexpr int ATT_THREADS = 128;
constexpr int ATT_CHUNK = 32;    // keys (fwd / dq) or queries (dkv) staged per iteration

template <class CT> __device__ __forceinline__ float neg_fill();
template <> __device__ __forceinline__ float neg_fill<float>() { return -FLT_MAX; }
template <> __device__ __forceinline__ float neg_fill<bf16>() { return -3.3895313892515355e38f; }  // finfo(bf16).min

struct AttnArgs {
  Segs q, kv;
  const float* kmask[NDIR];   // [B, Tk] key keep-mask of the direction's key modality, or null
  const void* qb;             // CT [MQ,256]   scaled queries
  const void* kvbuf;          // CT [MK, ldkv] ; K at col0 + h*32, V at col0 + 256 + h*32
  int ldkv, col0;
  void* o;                    // CT [MQ,256]
  float* ml;                  // fp32 [MQ, 8, 2]  (row max, 1/row sum)
  // backward
  const void* d_o;            // CT [MQ,256]
  void* dq;                   // CT [MQ,256]
  void* dkv;                  // CT [MK, ldkv]
  float* dvec;                // fp32 [MQ, 8]  sum_c dO*O
};

template <class CT>
__device__ __forceinline__ void load32(const CT* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = Vec4<CT>::ld(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
template <class CT>
__device__ __forceinline__ void store32(CT* p, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) Vec4<CT>::st(p + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
}
__device__ __forceinline__ float dot32(const float (&a)[32], const float* b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = *reinterpret_cast<const float4*>(b + 4 * i);
    s = fmaf(a[4 * i], t.x, s); s = fmaf(a[4 * i + 1], t.y, s);
    s = fmaf(a[4 * i + 2], t.z, s); s = fmaf(a[4 * i + 3], t.w, s);
  }
  return s;
}
__device__ __forceinline__ void axpy32(float (&acc)[32], float a, const float* b) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4 t = *reinterpret_cast<const float4*>(b + 4 * i);
    acc[4 * i] = fmaf(a, t.x, acc[4 * i]); acc[4 * i + 1] = fmaf(a, t.y, acc[4 * i + 1]);
    acc[4 * i + 2] = fmaf(a, t.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(a, t.w, acc[4 * i + 3]);
  }
}

// stage rows [j0, j0+n) x 256 columns (starting at column c0 of a CT matrix with leading dim ld)
// into fp32 shared memory [ATT_CHUNK][256]
template <class CT>
__device__ __forceinline__ void stage_rows(float* dst, const CT* src, size_t ld, int n) {
  for (int idx = threadIdx.x; idx < n * 64; idx += ATT_THREADS) {
    const int j = idx >> 6, c = (idx & 63) * 4;
    *reinterpret_cast<float4*>(dst + j * 256 + c) = Vec4<CT>::ld(src + (size_t)j * ld + c);
  }
}

// grid: (ceil(8*maxTq/128), B, 6)
template <class CT>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;                          // [32][256]
  float* Vs = smem + ATT_CHUNK * 256;        // [32][256]
  float* Ms = Vs + ATT_CHUNK * 256;          // [32] key keep flags
  const int d = blockIdx.z, b = blockIdx.y;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  if ((int)blockIdx.x * ATT_THREADS >= H * Tq) return;
  const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = i < H * Tq;
  const int h = active ? i / Tq : 0, t = active ? i % Tq : 0;
  const size_t qrow = (size_t)a.q.row0[d] + (size_t)qs_ + t;
  const CT* qb = reinterpret_cast<const CT*>(a.qb);
  const CT* kvb = reinterpret_cast<const CT*>(a.kvbuf) + ((size_t)a.kv.row0[d] + (size_t)b * Tk) * a.ldkv + a.col0;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
  const float NEG = neg_fill<CT>();
  float q[32], acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { q[c] = 0.f; acc[c] = 0.f; }
  if (active) load32<CT>(qb + qrow * D + h * HD, q);
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < Tk; j0 += ATT_CHUNK) {
    const int n = min(ATT_CHUNK, Tk - j0);
    __syncthreads();
    stage_rows<CT>(Ks, kvb + (size_t)j0 * a.ldkv, a.ldkv, n);
    stage_rows<CT>(Vs, kvb + (size_t)j0 * a.ldkv + D, a.ldkv, n);
    if (threadIdx.x < n) Ms[threadIdx.x] = km ? km[j0 + threadIdx.x] : 1.f;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      float s = round_ct<CT>(dot32(q, Ks + j * 256 + h * HD));
      if (Ms[j] < 0.5f) s = NEG;
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn);
      const float p = __expf(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] *= corr;
      axpy32(acc, p, Vs + j * 256 + h * HD);
      m = mn;
    }
  }
  if (active) {
    const float il = 1.0f / l;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] *= il;
    store32<CT>(reinterpret_cast<CT*>(a.o) + qrow * D + h * HD, acc);
    a.ml[(qrow * H + h) * 2] = m;
    a.ml[(qrow * H + h) * 2 + 1] = il;
  }
}

// dQ: same thread mapping as forward.  ds = p*(dp - D) for kept keys, 0 for padded keys
// (masked_fill blocks the gradient), dq = sum_j ds*k_j.
template <class CT>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dq_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = smem + ATT_CHUNK * 256;
  float* Ms = Vs + ATT_CHUNK * 256;
  const int d = blockIdx.z, b = blockIdx.y;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  if ((int)blockIdx.x * ATT_THREADS >= H * Tq) return;
  const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = i < H * Tq;
  const int h = active ? i / Tq : 0, t = active ? i % Tq : 0;
  const size_t qrow = (size_t)a.q.row0[d] + (size_t)qs_ + t;
  const CT* kvb = reinterpret_cast<const CT*>(a.kvbuf) + ((size_t)a.kv.row0[d] + (size_t)b * Tk) * a.ldkv + a.col0;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
  const float NEG = neg_fill<CT>();
  float q[32], dO[32], dq[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { q[c] = 0.f; dO[c] = 0.f; dq[c] = 0.f; }
  float m = 0.f, il = 0.f, Dv = 0.f;
  if (active) {
    load32<CT>(reinterpret_cast<const CT*>(a.qb) + qrow * D + h * HD, q);
    load32<CT>(reinterpret_cast<const CT*>(a.d_o) + qrow * D + h * HD, dO);
    float o[32];
    load32<CT>(reinterpret_cast<const CT*>(a.o) + qrow * D + h * HD, o);
#pragma unroll
    for (int c = 0; c < 32; ++c) Dv = fmaf(dO[c], o[c], Dv);
    m = a.ml[(qrow * H + h) * 2];
    il = a.ml[(qrow * H + h) * 2 + 1];
    a.dvec[qrow * H + h] = Dv;
  }
  for (int j0 = 0; j0 < Tk; j0 += ATT_CHUNK) {
    const int n = min(ATT_CHUNK, Tk - j0);
    __syncthreads();
    stage_rows<CT>(Ks, kvb + (size_t)j0 * a.ldkv, a.ldkv, n);
    stage_rows<CT>(Vs, kvb + (size_t)j0 * a.ldkv + D, a.ldkv, n);
    if (threadIdx.x < n) Ms[threadIdx.x] = km ? km[j0 + threadIdx.x] : 1.f;
    __syncthreads();
    for (int j = 0; j < n; ++j) {
      if (Ms[j] < 0.5f) continue;   // padded key: no gradient through the filled score
      const float s = round_ct<CT>(dot32(q, Ks + j * 256 + h * HD));
      const float p = __expf(s - m) * il;
      const float dp = dot32(dO, Vs + j * 256 + h * HD);
      axpy32(dq, p * (dp - Dv), Ks + j * 256 + h * HD);
    }
  }
  if (active) store32<CT>(reinterpret_cast<CT*>(a.dq) + qrow * D + h * HD, dq);
}

// dK/dV: one thread per (head, key); queries of the patient are staged in chunks.
// grid: (ceil(8*maxTk/128), B, 6)
template <class CT>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_dkv_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                           // [32][256]
  float* Gs = smem + ATT_CHUNK * 256;         // [32][256] dO
  float* St = Gs + ATT_CHUNK * 256;           // [32][8][3]: m, 1/l, D
  const int d = blockIdx.z, b = blockIdx.y;
  int qs_, Tq;
  seg_patient(a.q, d, b, qs_, Tq);     // this patient's (packed) query rows
  const int Tk = a.kv.T[d];
  if ((int)blockIdx.x * ATT_THREADS >= H * Tk) return;
  const int i = blockIdx.x * ATT_THREADS + threadIdx.x;
  const bool active = i < H * Tk;
  const int h = active ? i / Tk : 0, j = active ? i % Tk : 0;
  const size_t kvrow = (size_t)a.kv.row0[d] + (size_t)b * Tk + j;
  const CT* kvb = reinterpret_cast<const CT*>(a.kvbuf) + kvrow * a.ldkv + a.col0;
  const float* km = a.kmask[d] ? a.kmask[d] + (size_t)b * Tk : nullptr;
  const bool kept = active ? (km ? km[j] >= 0.5f : true) : false;
  const float NEG = neg_fill<CT>();
  float k[32], v[32], dk[32], dv[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
  if (active) { load32<CT>(kvb + h * HD, k); load32<CT>(kvb + D + h * HD, v); }
  const size_t q0 = (size_t)a.q.row0[d] + (size_t)qs_;
  const CT* qb = reinterpret_cast<const CT*>(a.qb) + q0 * D;
  const CT* gb = reinterpret_cast<const CT*>(a.d_o) + q0 * D;
  for (int i0 = 0; i0 < Tq; i0 += ATT_CHUNK) {
    const int n = min(ATT_CHUNK, Tq - i0);
    __syncthreads();
    stage_rows<CT>(Qs, qb + (size_t)i0 * D, D, n);
    stage_rows<CT>(Gs, gb + (size_t)i0 * D, D, n);
    for (int idx = threadIdx.x; idx < n * H; idx += ATT_THREADS) {
      const size_t r = (q0 + i0) * H + idx;
      St[idx * 3] = a.ml[r * 2]; St[idx * 3 + 1] = a.ml[r * 2 + 1]; St[idx * 3 + 2] = a.dvec[r];
    }
    __syncthreads();
    for (int ii = 0; ii < n; ++ii) {
      const float* st = St + (ii * H + h) * 3;
      float s = round_ct<CT>(dot32(k, Qs + ii * 256 + h * HD));
      if (!kept) s = NEG;
      const float p = __expf(s - st[0]) * st[1];
      axpy32(dv, p, Gs + ii * 256 + h * HD);
      if (kept) {
        const float dp = dot32(v, Gs + ii * 256 + h * HD);
        axpy32(dk, p * (dp - st[2]), Qs + ii * 256 + h * HD);
      }
    }
  }
  if (active) {
    CT* out = reinterpret_cast<CT*>(a.dkv) + kvrow * a.ldkv + a.col0;
    store32<CT>(out + h * HD, dk);
    store32<CT>(out + D + h * HD, dv);
  }
}

constexpr int ATT_SMEM_BYTES = (2 * ATT_CHUNK * 256 + ATT_CHUNK * H * 3 + 64) * 4;

}  // namespace mmr
