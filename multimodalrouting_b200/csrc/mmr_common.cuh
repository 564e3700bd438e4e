// Common device helpers for the B200 (sm_100a) route-fusion kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/mmr_b200.h"

namespace mmr {

using bf16 = __nv_bfloat16;

constexpr int D = MMR_D;          // model width
constexpr int H = MMR_HEADS;
constexpr int HD = MMR_HEAD_DIM;
constexpr int FF = MMR_FFN;
constexpr int NR = MMR_ROUTES;
constexpr int PC = MMR_PC_DIM;
constexpr int MC = MMR_MC_DIM;
constexpr int NDIR = 6;
constexpr int NMOD = 3;
constexpr float LN_EPS = 1e-5f;
constexpr float EMBED_SCALE = 16.0f;  // sqrt(256), transformer.py:33

// Direction d: query modality / key modality (0=L,1=N,2=I): LN, LI, NL, NI, IL, IN
__host__ __device__ inline int dir_qmod(int d) { return d >> 1; }
__host__ __device__ inline int dir_kmod(int d) {
  const int t[6] = {1, 2, 0, 2, 0, 1};
  return t[d];
}

// A set of up to 6 row segments laid out back to back (128-row aligned starts).
//
// Packed query rows (round 2): padded tokens of a query stream contribute nothing to the outputs or the gradients (every
// q-space tensor is multiplied by the query keep-mask, transformer.py:77-79,111-113), so the query row space holds only the
// VALID tokens of each patient, back to back.  Their number is data dependent and lives in device memory (one CUDA graph
// serves every batch): `nv[d]` = valid rows of segment d, `poff[d][b]` = first row of patient b inside the segment
// (`poff[d][B]` = nv[d]), `rowpat[d][j]` = patient of row j.  `rows[d]` / `T[d]` stay the static upper bounds (allocation, grid
// sizes).  Null pointers = the dense layout (row = b * T + t): the modality and key/value row spaces, and the query space when
// packing is off.  Contract for every q-space tensor: rows [nv, pad256(nv)) are zero, rows beyond are never touched.
struct Segs {
  int n;
  int row0[7];   // padded start row of each segment; row0[n] = total padded rows
  int rows[6];   // valid rows in the segment (upper bound when nv != nullptr)
  int T[6];      // tokens per patient in this segment (upper bound when poff != nullptr)
  const int* nv;          // device [6] or null
  const int* poff[6];     // device [B + 1] per segment or null
  const int* rowpat[6];   // device [rows] per segment or null
};

__device__ __forceinline__ int seg_rows(const Segs& s, int d) { return s.nv ? s.nv[d] : s.rows[d]; }
// rows [seg_rows, seg_rows_z) of a q-space tensor must be written as zeros (the tcgen05 tiles / CTA pairs read them)
__device__ __forceinline__ int seg_rows_z(const Segs& s, int d) {
  if (!s.nv) return s.row0[d + 1] - s.row0[d];
  const int z = (s.nv[d] + 255) & ~255, cap = s.row0[d + 1] - s.row0[d];
  return z < cap ? z : cap;
}
// first row (relative to the segment) and row count of patient b
__device__ __forceinline__ void seg_patient(const Segs& s, int d, int b, int& start, int& len) {
  if (s.poff[d]) { start = s.poff[d][b]; len = s.poff[d][b + 1] - start; }
  else { start = b * s.T[d]; len = s.T[d]; }
}
__device__ __forceinline__ int seg_row_patient(const Segs& s, int d, int local) {
  return s.rowpat[d] ? s.rowpat[d][local] : local / s.T[d];
}

__host__ __device__ inline int seg_of_row(const Segs& s, int row) {
  int d = 0;
#pragma unroll
  for (int i = 1; i < 6; ++i)
    if (i < s.n && row >= s.row0[i]) d = i;
  return d;
}

// ---- element access: CT in {float, bf16}, 4 consecutive elements at a time -------------------
template <class T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ float4 ld(const bf16* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
  }
  static __device__ __forceinline__ void st(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
    __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

// fp16 storage (saturating): 11-bit mantissa for the routing votes held in shared memory
template <> struct Vec4<__half> {
  static __device__ __forceinline__ float4 ld(const __half* p) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<__half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<__half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void st(__half* p, float4 v) {
    const float L = 65504.f;
    __half2 a = __floats2half2_rn(fminf(fmaxf(v.x, -L), L), fminf(fmaxf(v.y, -L), L));
    __half2 b = __floats2half2_rn(fminf(fmaxf(v.z, -L), L), fminf(fmaxf(v.w, -L), L));
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
  }
};

template <class T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// round-trip through the compute type (identity for fp32): mimics autocast rounding points
template <class T> __device__ __forceinline__ float round_ct(float v) { return to_f<T>(from_f<T>(v)); }

// Programmatic dependent launch: a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still draining; pdl_trigger() (first statement) lets the NEXT kernel do the same, pdl_wait() blocks
// until the predecessor grid has completed and its writes are visible -- nothing produced by an earlier kernel
// may be read, and nothing an earlier kernel reads may be written, before it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// A 256-wide fp32 row held by one warp: lane owns columns [4*lane, 4*lane+4) and [128+4*lane, ...).
struct Row8 {
  float v[8];
};
template <class T> __device__ __forceinline__ Row8 row_load(const T* rowptr, int lane) {
  Row8 r;
  float4 a = Vec4<T>::ld(rowptr + 4 * lane);
  float4 b = Vec4<T>::ld(rowptr + 128 + 4 * lane);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class T> __device__ __forceinline__ void row_store(T* rowptr, int lane, const Row8& r) {
  Vec4<T>::st(rowptr + 4 * lane, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
  Vec4<T>::st(rowptr + 128 + 4 * lane, make_float4(r.v[4], r.v[5], r.v[6], r.v[7]));
}
__device__ __forceinline__ int row_col(int lane, int i) { return (i < 4) ? 4 * lane + i : 128 + 4 * lane + (i - 4); }

// LayerNorm statistics of a 256-wide row distributed over a warp (two-pass, like ATen).
__device__ __forceinline__ void row_stats(const Row8& r, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r.v[i];
  mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d = r.v[i] - mean;
    q += d * d;
  }
  float var = warp_sum(q) * (1.0f / D);
  rstd = rsqrtf(var + LN_EPS);
}

}  // namespace mmr
