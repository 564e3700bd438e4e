// C ABI (include/mmr_b200.h) and host orchestration of the sm_100a route-fusion + routing path.
#include <stdio.h>
#include <stdlib.h>

#include <string>

#include "attention.cuh"
#include "attention_mma.cuh"
#include "epilogue.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"
#include "mmr_common.cuh"
#include "plan.cuh"
#include "routing.cuh"
#include "routing_split.cuh"
#include "rows.cuh"
#include "attention_tc.cuh"
#include "tail.cuh"
#include "loss.cuh"
#include "projector.cuh"
#include "producer.cuh"

using namespace mmr;

static thread_local std::string g_err;

// ---- instrumentation: launch counter + optional per-class CUDA-event timing (bench / profiling only)
#include <atomic>
#include <mutex>
#include <vector>
enum ProfClass { PC_GEMM_TC = 0, PC_WGRAD_TC = 1, PC_ATTN_FWD = 2, PC_ATTN_BWD = 3, PC_GEMM_SIMT = 4,
                 PC_ROUTING = 5, PC_FUSION_FWD = 6, PC_FUSION_BWD = 7, PC_N = 8 };
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
struct ProfRec { cudaEvent_t a, b; int cls; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof_recs;
struct ProfScope {
  cudaStream_t st; cudaEvent_t a = nullptr, b = nullptr; int cls; bool on;
  ProfScope(int c, cudaStream_t s) : st(s), cls(c), on(g_prof_on.load() != 0) {
    if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(b, st);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof_recs.push_back(ProfRec{a, b, cls});
    }
  }
};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(MMR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)
#define LAUNCH_OK(what)                                                                      \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return fail(MMR_ERR_CUDA, std::string("launch ") + what + ": " + cudaGetErrorString(_e)); \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
  } while (0)

// ------------------------------------------------------------------------------------------
// GEMM dispatch helpers.  One logical epilogue (EpiSpec) is lowered to the SIMT engine's element-wise
// functor or to the tcgen05 engine's store-only epilogue.
enum GOp { G_BIAS, G_BIAS_RELU, G_RELU_BWD, G_MASK, G_F32 };
struct EpiSpec {
  const float* bias = nullptr;      // G_BIAS, G_BIAS_RELU
  const float* rowmask = nullptr;   // G_MASK
  const void* relu_src = nullptr;   // G_RELU_BWD, SIMT engine: fc1 output (CT) [rows, ld_relu]
  int ld_relu = 0;
  uint32_t* bits = nullptr;         // tcgen05 engine: ReLU sign bits [rows, ld_bits] (written by G_BIAS_RELU)
  int ld_bits = 0;
  void* out = nullptr;
  int ldo = 0;
};

template <class CT, int GOP>
static int run_gemm(const Plan& P, const GemmProblem& g, const EpiSpec& sp, int a_rows, int b_rows,
                    cudaStream_t st, const char* what) {
  if (P.tc) {
    if (g.K % tc::BK != 0 || g.N % 256 != 0) return fail(MMR_ERR_UNSUPPORTED, std::string(what) + ": tcgen05 tile constraint");
    tc::TcEpi e; memset(&e, 0, sizeof(e));
    e.bias = sp.bias; e.rowmask = sp.rowmask; e.bits_in = sp.bits; e.bits_out = sp.bits; e.ld_bits = sp.ld_bits;
    e.out = sp.out; e.ldo = sp.ldo;
    constexpr int TOP = GOP == G_BIAS ? tc::TEPI_BIAS : GOP == G_BIAS_RELU ? tc::TEPI_BIAS_RELU_BITS
                      : GOP == G_RELU_BWD ? tc::TEPI_BITS_IN : GOP == G_MASK ? tc::TEPI_MASK : tc::TEPI_F32;
    ProfScope ps(PC_GEMM_TC, st);
    cudaError_t err = tc::launch_gemm_tc<TOP>(g, e, a_rows, b_rows, st);
    if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 gemm ") + what + ": " + cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    EpiParams e; memset(&e, 0, sizeof(e));
    e.bias = sp.bias; e.rowmask = sp.rowmask; e.aux = sp.relu_src; e.ldaux = sp.ld_relu; e.out = sp.out; e.ldo = sp.ldo;
    constexpr int SOP = GOP == G_BIAS ? EPI_BIAS : GOP == G_BIAS_RELU ? EPI_BIAS_RELU
                      : GOP == G_RELU_BWD ? EPI_RELUMASK : GOP == G_MASK ? EPI_MASK : EPI_STORE_F32;
    ProfScope ps(PC_GEMM_SIMT, st);
    launch_gemm_simt<CT, CT, SOP, CT>(g, e, st);
    LAUNCH_OK(what);
  }
  return MMR_OK;
}

// tcgen05 engine: bias gradients (column sums of dY) ride along in the weight-gradient kernel
static bool fuse_colsum(const Plan& P, WgradProblem& w, float* const* dbias) {
  if (!P.tc || getenv("MMR_NO_FUSED_COLSUM")) return false;
  for (int i = 0; i < w.segs.n; ++i) w.dbias[i] = dbias[i];
  w.colsum = 1;
  return true;
}

template <class CT>
static int run_wgrad(const Plan& P, const WgradProblem& w, int y_rows, int x_rows, cudaStream_t st, const char* what) {
  if (P.tc) {
    ProfScope ps(PC_WGRAD_TC, st);
    cudaError_t err = tc::launch_wgrad_tc(w, y_rows, x_rows, st);
    if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 wgrad ") + what + ": " + cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    ProfScope ps(PC_GEMM_SIMT, st);
    launch_wgrad_simt<CT, CT>(w, st);
    LAUNCH_OK(what);
  }
  return MMR_OK;
}

constexpr int AHG = 2;   // heads per attention CTA (attention_mma.cuh)
// bf16 path: tensor-core attention (attention_mma.cuh); fp32 parity path and MMR_ATTN=simt: SIMT kernels
template <class CT> static bool mma_attention() { return false; }
template <> bool mma_attention<bf16>() {
  const char* e = getenv("MMR_ATTN");   // read per call so tests can switch engines inside one process
  return !(e && !strcmp(e, "simt"));
}

// MMR_ATTN=tc: tcgen05 / TMEM / TMA attention forward (attention_tc.cuh), the long-sequence engine; the backward stays on
// the mma.sync kernels, which consume the same (o, ml) outputs
template <class CT> static bool tc_attention() { return false; }
template <> bool tc_attention<bf16>() {
  const char* e = getenv("MMR_ATTN");
  return e && !strcmp(e, "tc");
}

// launches a kernel that implements the pdl_trigger / pdl_wait protocol (programmatic dependent launch)
template <class K, class A>
static void launch_k(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const A& a) {
  if (tc::pdl_enabled()) launch_pdl(kern, grid, block, smem, st, a);
  else kern<<<grid, block, smem, st>>>(a);
}

static Segs single_seg(int rows, int T) {
  Segs s;
  memset(&s, 0, sizeof(s));
  s.n = 1; s.row0[0] = 0; s.row0[1] = rows; s.rows[0] = rows; s.T[0] = T;
  return s;
}

// plain fp32 GEMM  C[M,N] (ldc) = A[M,K] (lda) * W^T (+bias), W [N,K] (TRANSB=false) or [K,N] (true)
template <bool TRANSB>
static int fp32_linear(const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc,
                       int M, int N, int K, cudaStream_t st, const char* what, int tf32 = 0) {
  GemmProblem g;
  memset(&g, 0, sizeof(g));
  g.segs = single_seg(M, 1);
  g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = W; g.ldb = ldw; g.tf32 = tf32;
  EpiParams e;
  memset(&e, 0, sizeof(e));
  e.bias = bias; e.out = C; e.ldo = ldc;
  launch_gemm_simt<float, float, EPI_BIAS_F32, float, TRANSB>(g, e, st);
  LAUNCH_OK(what);
  return MMR_OK;
}

static void fp32_problem(GemmProblem* g, EpiParams* e, const float* A, int lda, const float* W, int ldw, const float* bias,
                         float* C, int ldc, int M, int N, int K, int tf32 = 0) {
  memset(g, 0, sizeof(*g));
  g->segs = single_seg(M, 1);
  g->N = N; g->K = K; g->A = A; g->lda = lda; g->B = W; g->ldb = ldw; g->tf32 = tf32;
  memset(e, 0, sizeof(*e));
  e->bias = bias; e->out = C; e->ldo = ldc;
}

static int fp32_wgrad(const float* dY, int ldy, const float* X, int ldx, float* out, int ldo, int rows, int M, int N,
                      cudaStream_t st, const char* what, int tf32 = 0) {
  if (out == nullptr) return MMR_OK;
  WgradProblem w;
  memset(&w, 0, sizeof(w));
  w.segs = single_seg(rows, 1);
  w.dY = dY; w.ldy = ldy; w.X = X; w.ldx = ldx; w.M = M; w.N = N; w.out[0] = out; w.ldo = ldo; w.tf32 = tf32;
  launch_wgrad_simt<float, float>(w, st);
  LAUNCH_OK(what);
  return MMR_OK;
}

template <class T>
static int run_colsum(const Segs& segs, const void* src, int ld, int col0, int ncols, float* const* out, float scale,
                      cudaStream_t st, const char* what) {
  ColsumArgs c;
  memset(&c, 0, sizeof(c));
  c.segs = segs; c.src = src; c.ld = ld; c.col0 = col0; c.ncols = ncols; c.scale = scale;
  bool any = false;
  for (int i = 0; i < segs.n; ++i) { c.out[i] = out[i]; any = any || out[i] != nullptr; }
  if (!any) return MMR_OK;
  const int total = segs.row0[segs.n];
  dim3 grid((ncols + 255) / 256, (total + 127) / 128);
  colsum_kernel<T><<<grid, 256, 0, st>>>(c);
  LAUNCH_OK(what);
  return MMR_OK;
}

static bool has_pad(const Segs& s) {
  for (int i = 0; i < s.n; ++i)
    if (s.row0[i + 1] - s.row0[i] != s.rows[i]) return true;
  return false;
}
static int zero_pad(const Segs& s, void* buf, size_t ld_bytes, cudaStream_t st) {
  if (s.nv == nullptr && !has_pad(s)) return MMR_OK;     // packed query rows: the pad zone is data dependent, always run
  dim3 grid(255, s.n);
  zero_pad_rows_kernel<<<grid, 128, 0, st>>>(s, reinterpret_cast<uint8_t*>(buf), ld_bytes);
  LAUNCH_OK("zero_pad_rows");
  return MMR_OK;
}

// ------------------------------------------------------------------------------------------
template <class CT>
static int pack_weights(const Plan& P, const void* const* prm, uint8_t* packed, cudaStream_t st) {
  const ParamIndex ix{P.L};
  const int L = P.L;
  const float scaling = 1.0f / sqrtf((float)HD);
  auto f = [&](int i) { return reinterpret_cast<const float*>(prm[i]); };
  CT* wq = reinterpret_cast<CT*>(packed + P.o_wq);   CT* wqT = reinterpret_cast<CT*>(packed + P.o_wqT);
  CT* wo = reinterpret_cast<CT*>(packed + P.o_wo);   CT* woT = reinterpret_cast<CT*>(packed + P.o_woT);
  CT* w1 = reinterpret_cast<CT*>(packed + P.o_w1);   CT* w1T = reinterpret_cast<CT*>(packed + P.o_w1T);
  CT* w2 = reinterpret_cast<CT*>(packed + P.o_w2);   CT* w2T = reinterpret_cast<CT*>(packed + P.o_w2T);
  CT* wkv = reinterpret_cast<CT*>(packed + P.o_wkv); CT* wkvT = reinterpret_cast<CT*>(packed + P.o_wkvT);
  float* bq = reinterpret_cast<float*>(packed + P.o_bq);  float* bo = reinterpret_cast<float*>(packed + P.o_bo);
  float* b1 = reinterpret_cast<float*>(packed + P.o_b1);  float* b2 = reinterpret_cast<float*>(packed + P.o_b2);
  float* bkv = reinterpret_cast<float*>(packed + P.o_bkv);
  PackJobs pj; memset(&pj, 0, sizeof(pj));
  BiasJobs bj; memset(&bj, 0, sizeof(bj));
  auto flush = [&]() -> int {
    if (pj.n == 0) return MMR_OK;
    pack_kernel<CT><<<dim3((FF / 32) * (D / 32), pj.n), 256, 0, st>>>(pj);
    LAUNCH_OK("pack_kernel");
    bias_fold_kernel<<<dim3(FF / 8, bj.n), 256, 0, st>>>(bj);
    LAUNCH_OK("bias_fold_kernel");
    pj.n = 0; bj.n = 0;
    return MMR_OK;
  };
  for (int d = 0; d < NDIR; ++d) {
    for (int l = 0; l < L; ++l) {
      if (pj.n + 5 > 60) { int rc = flush(); if (rc) return rc; }
      {
        const size_t ld_ = (size_t)l * 6 + d;
        const float* win = f(ix.layer(d, l, 0));
        const float* bin = f(ix.layer(d, l, 1));
        const float* g0 = f(ix.layer(d, l, 8));
        const float* be0 = f(ix.layer(d, l, 9));
        PackJob* j = &pj.j[pj.n++];
        *j = PackJob{win, D, D, D, nullptr, scaling, wq + ld_ * D * D, D, wqT + ld_ * D * D, D};
        j = &pj.j[pj.n++];
        *j = PackJob{win + (size_t)D * D, 2 * D, D, D, g0, 1.0f, wkv + ((size_t)d * L + l) * 2 * D * D, D,
                     wkvT + (size_t)d * D * (L * 2 * D) + (size_t)l * 2 * D, L * 2 * D};
        j = &pj.j[pj.n++];
        *j = PackJob{f(ix.layer(d, l, 2)), D, D, D, nullptr, 1.0f, wo + ld_ * D * D, D, woT + ld_ * D * D, D};
        j = &pj.j[pj.n++];
        *j = PackJob{f(ix.layer(d, l, 4)), FF, D, D, nullptr, 1.0f, w1 + ld_ * FF * D, D, w1T + ld_ * D * FF, FF};
        j = &pj.j[pj.n++];
        *j = PackJob{f(ix.layer(d, l, 6)), D, FF, FF, nullptr, 1.0f, w2 + ld_ * D * FF, FF, w2T + ld_ * FF * D, D};
        BiasJob* b = &bj.j[bj.n++];
        *b = BiasJob{nullptr, 0, bin, nullptr, scaling, bq + ld_ * D, D};
        b = &bj.j[bj.n++];
        *b = BiasJob{win + (size_t)D * D, D, bin + D, be0, 1.0f, bkv + ((size_t)d * L + l) * 2 * D, 2 * D};
        b = &bj.j[bj.n++];
        *b = BiasJob{nullptr, 0, f(ix.layer(d, l, 3)), nullptr, 1.0f, bo + ld_ * D, D};
        b = &bj.j[bj.n++];
        *b = BiasJob{nullptr, 0, f(ix.layer(d, l, 5)), nullptr, 1.0f, b1 + ld_ * FF, FF};
        b = &bj.j[bj.n++];
        *b = BiasJob{nullptr, 0, f(ix.layer(d, l, 7)), nullptr, 1.0f, b2 + ld_ * D, D};
      }
    }
  }
  {
    int rc = flush();
    if (rc) return rc;
  }
  return MMR_OK;
}

// The query row space with its device-side row plan (Segs in mmr_common.cuh; the plan lives in `saved`).
static Segs query_segs(const Plan& P, const uint8_t* saved) {
  Segs Q = P.q;
  const int* plan = reinterpret_cast<const int*>(saved + P.s_plan);
  Q.nv = plan;
  for (int d = 0; d < NDIR; ++d) {
    Q.poff[d] = plan + P.p_poff[dir_qmod(d)];
    Q.rowpat[d] = plan + P.p_rowpat[dir_qmod(d)];
  }
  return Q;
}
// MMR_VARLEN=0 keeps the dense query layout (every token has a row); the tcgen05 attention engine needs it (its 3-D
// tensor maps address [patient][token][column]).
static bool pack_query_rows(bool tc_attn) {
  const char* e = getenv("MMR_VARLEN");
  return !(e && e[0] == '0') && !tc_attn;
}

template <class CT>
static int fusion_fwd(const Plan& P, const void* const* prm, const float* const x[3], const float* const mask[3],
                      const float* pos, uint8_t* packed, uint8_t* saved, uint8_t* scratch, float* routes,
                      cudaStream_t st, bool do_pack) {
  const ParamIndex ix{P.L};
  // reduced-precision mode: the small fp32 GEMMs (Conv1d input projections, pair / trimodal composition) contract on tf32
  // tensor cores (the reference runs these nn.Linear / Conv1d in bf16 under autocast); MMR_TF32=0 keeps the FMA loops
  static const int tf_on = [] { const char* e = getenv("MMR_TF32"); return (e && atoi(e) == 0) ? 0 : 1; }();
  const int TF32 = std::is_same<CT, bf16>::value ? tf_on : 0;
  const int L = P.L, B = P.B;
  auto f = [&](int i) { return reinterpret_cast<const float*>(prm[i]); };
  int rc = do_pack ? pack_weights<CT>(P, prm, packed, st) : MMR_OK;   // else: `packed` holds mmr_fusion_pack_weights' output
  if (rc) return rc;
  const Segs Q = query_segs(P, saved);
  {  // row plan of the (packed) query space
    RowPlanArgs a; memset(&a, 0, sizeof(a));
    int* plan = reinterpret_cast<int*>(saved + P.s_plan);
    a.B = B; a.pack = pack_query_rows(tc_attention<CT>()) ? 1 : 0; a.nv = plan;
    for (int m = 0; m < NMOD; ++m) {
      a.mask[m] = mask[m]; a.T[m] = P.T[m];
      a.poff[m] = plan + P.p_poff[m]; a.tokrow[m] = plan + P.p_tokrow[m]; a.rowpat[m] = plan + P.p_rowpat[m];
    }
    const dim3 gp((B + 31) / 32, NMOD);
    rowplan_count_kernel<<<gp, 1024, 0, st>>>(a);
    LAUNCH_OK("rowplan_count");
    rowplan_scan_kernel<<<NMOD, 1024, 0, st>>>(a);
    LAUNCH_OK("rowplan_scan");
    rowplan_fill_kernel<<<gp, 1024, 0, st>>>(a);
    LAUNCH_OK("rowplan_fill");
  }

  float* fp = reinterpret_cast<float*>(scratch + P.f_p);
  float* fy = reinterpret_cast<float*>(scratch + P.f_y);
  float* fu = reinterpret_cast<float*>(scratch + P.f_u);
  CT* delta = reinterpret_cast<CT*>(scratch + P.f_delta);
  auto bits = [&](int l) { return reinterpret_cast<uint32_t*>(saved + P.s_bits + (size_t)l * P.l_bits); };
  CT* xh = reinterpret_cast<CT*>(saved + P.s_xh);
  float* maskq = reinterpret_cast<float*>(saved + P.s_maskq);
  auto xin = [&](int l) { return reinterpret_cast<float*>(saved + P.s_xin + (size_t)l * P.l_xin); };
  auto x1 = [&](int l) { return reinterpret_cast<float*>(saved + P.s_x1 + (size_t)l * P.l_xin); };
  auto stat0 = [&](int l) { return reinterpret_cast<float*>(saved + P.s_stat0 + (size_t)l * P.l_stat); };
  auto stat1 = [&](int l) { return reinterpret_cast<float*>(saved + P.s_stat1 + (size_t)l * P.l_stat); };
  auto h0 = [&](int l) { return reinterpret_cast<CT*>(saved + P.s_h0 + (size_t)l * P.l_ct256); };
  auto h1 = [&](int l) { return reinterpret_cast<CT*>(saved + P.s_h1 + (size_t)l * P.l_ct256); };
  auto qb = [&](int l) { return reinterpret_cast<CT*>(saved + P.s_qb + (size_t)l * P.l_ct256); };
  auto ob = [&](int l) { return reinterpret_cast<CT*>(saved + P.s_o + (size_t)l * P.l_ct256); };
  auto ml = [&](int l) { return reinterpret_cast<float*>(saved + P.s_ml + (size_t)l * P.l_ml); };
  auto ff = [&](int l) { return reinterpret_cast<CT*>(saved + P.s_f + (size_t)l * P.l_f); };
  CT* kv = reinterpret_cast<CT*>(saved + P.s_kv);
  float* ecat = reinterpret_cast<float*>(saved + P.s_epair);
  float* zcat = reinterpret_cast<float*>(saved + P.s_routes);
  float* cnt = reinterpret_cast<float*>(saved + P.s_cnt);
  const int ldkv = L * 2 * D;

  // 1. optional Conv1d(k=1) projections (mult_model.py:134-136)
  const float* src[3];
  for (int m = 0; m < NMOD; ++m) {
    if (P.din[m] == D) { src[m] = x[m]; continue; }
    float* dst = fp + (size_t)P.mod.row0[m] * D;
    rc = fp32_linear<false>(x[m], P.din[m], f(ix.proj(m)), P.din[m], nullptr, dst, D, P.mod.rows[m], D, P.din[m], st, "proj", TF32);
    if (rc) return rc;
    src[m] = dst;
  }
  // 2. embedding, unimodal encoders, normalised K/V stream, layer-0 query streams
  {
    EmbedArgs a; memset(&a, 0, sizeof(a));
    a.mod = P.mod; a.q = Q; a.pos = pos;
    for (int m = 0; m < NMOD; ++m) {
      a.src[m] = src[m]; a.mask[m] = mask[m];
      a.uni_g[m] = f(ix.uni_ln(m, 0)); a.uni_b[m] = f(ix.uni_ln(m, 1));
    }
    for (int d = 0; d < NDIR; ++d) { a.ln0_g[d] = f(ix.layer(d, 0, 8)); a.ln0_b[d] = f(ix.layer(d, 0, 9)); }
    a.xh = xh; a.rstd_e = reinterpret_cast<float*>(saved + P.s_rstd_e); a.u = fu;
    a.xin0 = xin(0); a.h0 = h0(0); a.stat0 = stat0(0); a.maskq = maskq;
    for (int m = 0; m < NMOD; ++m) a.tokrow[m] = reinterpret_cast<const int*>(saved + P.s_plan) + P.p_tokrow[m];
    embed_fwd_kernel<CT><<<P.MM / ROWS_PER_BLOCK, 256, 0, st>>>(a);
    LAUNCH_OK("embed_fwd");
    rc = zero_pad(Q, h0(0), (size_t)D * sizeof(CT), st);     // layer-0 Q-projection operand: rows [nv, pad) must be zero
    if (rc) return rc;
  }
  // 3. K/V projections of every layer at once (LN0 affine folded into the weights)
  {
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = P.kv;
    for (int d = 0; d < NDIR; ++d) { g.a_row0[d] = P.mod.row0[dir_kmod(d)]; g.b_row0[d] = d * ldkv; }
    g.N = ldkv; g.K = D; g.A = xh; g.lda = D; g.B = packed + P.o_wkv; g.ldb = D;
    EpiSpec e;
    e.bias = reinterpret_cast<const float*>(packed + P.o_bkv); e.out = kv; e.ldo = ldkv;
    rc = run_gemm<CT, G_BIAS>(P, g, e, P.MM, NDIR * ldkv, st, "kv_proj");
    if (rc) return rc;
  }
  const float* kmask[NDIR];
  for (int d = 0; d < NDIR; ++d) kmask[d] = mask[dir_kmod(d)];
  int maxTq = 0;
  for (int d = 0; d < NDIR; ++d) maxTq = maxTq > Q.T[d] ? maxTq : Q.T[d];
  CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(amma::attn_fwd_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::fwd_smem<AHG>()));
  CUDA_OK(cudaFuncSetAttribute(amma::attn_fwd_single_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::fwd_smem<AHG>()));

  auto q_problem = [&](const void* A, int lda, const void* Bw, int ldb, int nrows_b, int N, int K, int l) {
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = Q;
    for (int d = 0; d < NDIR; ++d) { g.a_row0[d] = Q.row0[d]; g.b_row0[d] = (l * 6 + d) * nrows_b; }
    g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = Bw; g.ldb = ldb;
    return g;
  };

  for (int l = 0; l < L; ++l) {
    {  // Q projection (scaled)
      GemmProblem g = q_problem(h0(l), D, packed + P.o_wq, D, D, D, D, l);
      EpiSpec e;
      e.bias = reinterpret_cast<const float*>(packed + P.o_bq); e.out = qb(l); e.ldo = D;
      rc = run_gemm<CT, G_BIAS>(P, g, e, P.MQ, L * 6 * D, st, "q_proj");
      if (rc) return rc;
    }
    {  // attention
      AttnArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.kv = P.kv;
      for (int d = 0; d < NDIR; ++d) a.kmask[d] = kmask[d];
      a.qb = qb(l); a.kvbuf = kv; a.ldkv = ldkv; a.col0 = l * 2 * D; a.o = ob(l); a.ml = ml(l);
      bool pad_done = false;
      {
        ProfScope ps(PC_ATTN_FWD, st);
        if (tc_attention<CT>()) {
          cudaError_t ce = atc::launch_attn_fwd_tc(a, B, maxTq, st);
          if (ce != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("launch attn_fwd_tc: ") + cudaGetErrorString(ce));
        } else if (mma_attention<CT>()) {
          dim3 grid(amma::Cfg<AHG>::NHG * ((maxTq + amma::RC - 1) / amma::RC), B, NDIR);
          int maxTk = 0;
          for (int d = 0; d < NDIR; ++d) maxTk = maxTk > P.kv.T[d] ? maxTk : P.kv.T[d];
          if (maxTk <= amma::RC && !getenv("MMR_ATTN_FWD_GENERAL")) {
            amma::attn_fwd_single_kernel<AHG><<<grid, amma::Cfg<AHG>::THREADS, amma::fwd_smem<AHG>(), st>>>(a);
            pad_done = true;        // the kernel clears rows [nv, pad256(nv)) of its output itself
          } else
            launch_k(amma::attn_fwd_kernel<AHG>, grid, dim3(amma::Cfg<AHG>::THREADS), amma::fwd_smem<AHG>(), st, a);
        } else {
          dim3 grid((H * maxTq + ATT_THREADS - 1) / ATT_THREADS, B, NDIR);
          attn_fwd_kernel<CT><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
        }
      }
      LAUNCH_OK("attn_fwd");
      if (!pad_done) {
        rc = zero_pad(Q, ob(l), (size_t)D * sizeof(CT), st);
        if (rc) return rc;
      }
    }
    {  // out projection (+bias); the residual add + mask is fused into the LN1 kernel below
      GemmProblem g = q_problem(ob(l), D, packed + P.o_wo, D, D, D, D, l);
      EpiSpec e;
      e.bias = reinterpret_cast<const float*>(packed + P.o_bo); e.out = delta; e.ldo = D;
      rc = run_gemm<CT, G_BIAS>(P, g, e, P.MQ, L * 6 * D, st, "out_proj");
      if (rc) return rc;
    }
    {  // x1 = (x + attn)*mask ; h1 = LN1(x1)*mask
      LnFwdArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.x = xin(l); a.delta = delta; a.x_out = x1(l); a.maskq = maskq; a.out = h1(l); a.stat = stat1(l);
      for (int d = 0; d < NDIR; ++d) { a.gamma[d] = f(ix.layer(d, l, 10)); a.beta[d] = f(ix.layer(d, l, 11)); }
      launch_k(ln_rows_fwd_kernel<CT, CT>, dim3(P.MQ / ROWS_PER_BLOCK), dim3(256), 0, st, a);
      LAUNCH_OK("ln1_fwd");
    }
    {
      {  // fc1 + relu
        GemmProblem g = q_problem(h1(l), D, packed + P.o_w1, D, FF, FF, D, l);
        EpiSpec e;
        e.bias = reinterpret_cast<const float*>(packed + P.o_b1); e.out = ff(l); e.ldo = FF;
        e.bits = bits(l); e.ld_bits = FF / 32;
        rc = run_gemm<CT, G_BIAS_RELU>(P, g, e, P.MQ, L * 6 * FF, st, "fc1");
        if (rc) return rc;
      }
      {  // fc2 (+bias); residual add + mask fused into the following LayerNorm kernel
        GemmProblem g = q_problem(ff(l), FF, packed + P.o_w2, FF, D, D, FF, l);
        EpiSpec e;
        e.bias = reinterpret_cast<const float*>(packed + P.o_b2); e.out = delta; e.ldo = D;
        rc = run_gemm<CT, G_BIAS>(P, g, e, P.MQ, L * 6 * D, st, "fc2");
        if (rc) return rc;
      }
    }
    if (l + 1 < L) {  // x_{l+1} = (x1 + ffn)*mask ; h0 = LN0_{l+1}(x_{l+1})*mask
      LnFwdArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.x = x1(l); a.delta = delta; a.x_out = xin(l + 1); a.maskq = maskq; a.out = h0(l + 1); a.stat = stat0(l + 1);
      for (int d = 0; d < NDIR; ++d) { a.gamma[d] = f(ix.layer(d, l + 1, 8)); a.beta[d] = f(ix.layer(d, l + 1, 9)); }
      launch_k(ln_rows_fwd_kernel<CT, CT>, dim3(P.MQ / ROWS_PER_BLOCK), dim3(256), 0, st, a);
      LAUNCH_OK("ln0_fwd");
    }
  }
  {  // x_L = (x1 + ffn)*mask, then the encoder-final LayerNorm (transformer.py:108-113)
    LnFwdArgs a; memset(&a, 0, sizeof(a));
    a.q = Q; a.x = x1(L - 1); a.delta = delta; a.x_out = xin(L); a.maskq = maskq; a.out = fy;
    a.stat = reinterpret_cast<float*>(saved + P.s_statf);
    for (int d = 0; d < NDIR; ++d) { a.gamma[d] = f(ix.enc_ln(d, 0)); a.beta[d] = f(ix.enc_ln(d, 1)); }
    launch_k(ln_rows_fwd_kernel<float, CT>, dim3(P.MQ / ROWS_PER_BLOCK), dim3(256), 0, st, a);
    LAUNCH_OK("lnf_fwd");
  }
  {  // masked-mean pooling of the 9 uni/bi-modal routes
    PoolArgs a; memset(&a, 0, sizeof(a));
    a.mod = P.mod; a.q = Q; a.u = fu; a.y = fy; a.routes = routes; a.zcat = zcat; a.cnt = cnt; a.B = B;
    a.maskq = maskq;
    for (int m = 0; m < NMOD; ++m) a.mask[m] = mask[m];
    pool_fwd_kernel<<<dim3(B, 9), 256, 0, st>>>(a);
    LAUNCH_OK("pool_fwd");
  }
  // pair projections and the trimodal composition (mult_model.py:174-178)
  {
    MultiGemm mg; mg.n = 3;     // the three pair projections in one launch
    for (int p = 0; p < 3; ++p)
      fp32_problem(&mg.g[p], &mg.e[p], zcat + (size_t)p * B * 512, 512, f(ix.pair(p, 0)), 512, f(ix.pair(p, 1)), ecat + p * D,
                   3 * D, B, D, 512, TF32);
    launch_gemm_simt_multi<float, float, EPI_BIAS_F32, float, false>(mg, st);
    LAUNCH_OK("pair_proj");
  }
  rc = fp32_linear<false>(ecat, 3 * D, f(ix.final_lni(0)), 3 * D, f(ix.final_lni(1)), routes + (size_t)9 * B * D, D, B, D,
                          3 * D, st, "final_lni", TF32);
  return rc;
}

template <class CT>
static int fusion_bwd(const Plan& P, const void* const* prm, const float* const x[3], const float* const mask[3],
                      const uint8_t* packed, const uint8_t* saved, uint8_t* scratch, const float* d_routes,
                      void* const* grads, float* const dx[3], cudaStream_t st, void* const* layer_events = nullptr,
                      cudaStream_t side = nullptr, void* const* sync_ev = nullptr) {
  const ParamIndex ix{P.L};
  static const int tf_on = [] { const char* e = getenv("MMR_TF32"); return (e && atoi(e) == 0) ? 0 : 1; }();
  const int TF32 = std::is_same<CT, bf16>::value ? tf_on : 0;
  // Optional second stream for the weight-gradient kernels.  They depend only on a layer's upstream gradient and saved
  // activations, never feed the data-gradient chain, and are L2/HBM-latency bound, so they can run next to the
  // attention-backward and LayerNorm kernels of the chain.  Nine caller-owned events order the two streams:
  // main -> side when an operand is ready, side -> main before the chain overwrites an operand the side stream reads.
  enum { E_GC0 = 0, E_DF, E_GC1, E_DQ, S_DW2, S_DW1, S_DWO, S_DWQ, S_JOIN, N_SYNC_EV };
  static_assert(N_SYNC_EV == MMR_BWD_SYNC_EVENTS, "header constant out of date");
  const bool two = side != nullptr && sync_ev != nullptr && layer_events == nullptr;
  const cudaStream_t ws = two ? side : st;
  auto sev = [&](int i) { return reinterpret_cast<cudaEvent_t>(sync_ev[i]); };
  auto main_to_side = [&](int e) -> int {      // side stream continues once everything issued on main so far is done
    if (!two) return MMR_OK;
    CUDA_OK(cudaEventRecord(sev(e), st));
    CUDA_OK(cudaStreamWaitEvent(side, sev(e), 0));
    return MMR_OK;
  };
  auto side_record = [&](int e) -> int {
    if (!two) return MMR_OK;
    CUDA_OK(cudaEventRecord(sev(e), side));
    return MMR_OK;
  };
  auto main_wait = [&](int e) -> int {
    if (!two) return MMR_OK;
    CUDA_OK(cudaStreamWaitEvent(st, sev(e), 0));
    return MMR_OK;
  };
  bool have_dw1 = false, have_dwq = false, have_dwo = false;
  const int L = P.L, B = P.B;
  const Segs Q = query_segs(P, saved);
  auto f = [&](int i) { return reinterpret_cast<const float*>(prm[i]); };
  auto gr = [&](int i) { return reinterpret_cast<float*>(grads[i]); };
  int rc;
  CUDA_OK(cudaMemsetAsync(scratch + P.b_zero_begin, 0, P.b_zero_end - P.b_zero_begin, st));

  const CT* xh = reinterpret_cast<const CT*>(saved + P.s_xh);
  const float* maskq = reinterpret_cast<const float*>(saved + P.s_maskq);
  auto xin = [&](int l) { return reinterpret_cast<const float*>(saved + P.s_xin + (size_t)l * P.l_xin); };
  auto x1 = [&](int l) { return reinterpret_cast<const float*>(saved + P.s_x1 + (size_t)l * P.l_xin); };
  auto stat0 = [&](int l) { return reinterpret_cast<const float*>(saved + P.s_stat0 + (size_t)l * P.l_stat); };
  auto stat1 = [&](int l) { return reinterpret_cast<const float*>(saved + P.s_stat1 + (size_t)l * P.l_stat); };
  auto h0 = [&](int l) { return reinterpret_cast<const CT*>(saved + P.s_h0 + (size_t)l * P.l_ct256); };
  auto h1 = [&](int l) { return reinterpret_cast<const CT*>(saved + P.s_h1 + (size_t)l * P.l_ct256); };
  auto qb = [&](int l) { return reinterpret_cast<const CT*>(saved + P.s_qb + (size_t)l * P.l_ct256); };
  auto ob = [&](int l) { return reinterpret_cast<const CT*>(saved + P.s_o + (size_t)l * P.l_ct256); };
  auto ml = [&](int l) { return reinterpret_cast<const float*>(saved + P.s_ml + (size_t)l * P.l_ml); };
  auto ff = [&](int l) { return reinterpret_cast<const CT*>(saved + P.s_f + (size_t)l * P.l_f); };
  const CT* kv = reinterpret_cast<const CT*>(saved + P.s_kv);
  const float* ecat = reinterpret_cast<const float*>(saved + P.s_epair);
  const float* zcat = reinterpret_cast<const float*>(saved + P.s_routes);
  const float* cnt = reinterpret_cast<const float*>(saved + P.s_cnt);
  const int ldkv = L * 2 * D;

  float* g_a = reinterpret_cast<float*>(scratch + P.b_g);
  float* g_b = reinterpret_cast<float*>(scratch + P.b_g1);
  CT* gcbuf[2] = {reinterpret_cast<CT*>(scratch + P.b_gc), reinterpret_cast<CT*>(scratch + P.b_gc2)};
  CT* gc = gcbuf[0];   // bf16/CT copy of the current residual gradient; alternates between the two buffers
  int gci = 0;
  CT* dF = reinterpret_cast<CT*>(scratch + P.b_df);
  CT* dH = reinterpret_cast<CT*>(scratch + P.b_dh);
  CT* dO = reinterpret_cast<CT*>(scratch + P.b_do);
  CT* dQ = reinterpret_cast<CT*>(scratch + P.b_dq);
  CT* dKV = reinterpret_cast<CT*>(scratch + P.b_dkv);
  float* dxh = reinterpret_cast<float*>(scratch + P.b_dxh);
  float* dp = reinterpret_cast<float*>(scratch + P.b_dp);
  float* dwq = reinterpret_cast<float*>(scratch + P.b_dwq);
  float* dwkv = reinterpret_cast<float*>(scratch + P.b_dwkv);
  float* dbq = reinterpret_cast<float*>(scratch + P.b_dbq);
  float* dbkv = reinterpret_cast<float*>(scratch + P.b_dbkv);
  float* decat = reinterpret_cast<float*>(scratch + P.b_depair);   // [B,768]
  float* dzcat = reinterpret_cast<float*>(scratch + P.b_dzcat);    // [3,B,512]
  float* dvec = reinterpret_cast<float*>(scratch + P.b_dvec);      // [MQ,8]

  // ---- trimodal + pair projections (mult_model.py:174-178) ----
  const float* dz_lni = d_routes + (size_t)9 * B * D;
  // The weight / bias gradients of this tail (4 launches, ~75 us at B = 512) are off the data-gradient chain: they go to the
  // side stream, which is idle until the last layer's gradients exist.  E_GC0 / E_DF are re-recorded further down.
  rc = main_to_side(E_GC0);   // fork (after the memset of the accumulators)
  if (rc) return rc;
  rc = fp32_linear<true>(dz_lni, D, f(ix.final_lni(0)), 3 * D, nullptr, decat, 3 * D, B, 3 * D, D, st, "d_final_lni", TF32);
  if (rc) return rc;
  rc = main_to_side(E_DF);    // d(ecat) is ready
  if (rc) return rc;
  rc = fp32_wgrad(dz_lni, D, ecat, 3 * D, gr(ix.final_lni(0)), 3 * D, B, D, 3 * D, ws, "w_final_lni", TF32);
  if (rc) return rc;
  {
    Segs sb = single_seg(B, 1);
    float* o1[6] = {gr(ix.final_lni(1)), nullptr, nullptr, nullptr, nullptr, nullptr};
    rc = run_colsum<float>(sb, dz_lni, D, 0, D, o1, 1.0f, ws, "b_final_lni");
    if (rc) return rc;
    {
      MultiGemm mg; mg.n = 3;   // d(zcat_p) = d(e_p) W_p for the three pairs in one launch
      for (int p = 0; p < 3; ++p)
        fp32_problem(&mg.g[p], &mg.e[p], decat + p * D, 3 * D, f(ix.pair(p, 0)), 512, nullptr, dzcat + (size_t)p * B * 512, 512, B,
                     512, D, TF32);
      launch_gemm_simt_multi<float, float, EPI_BIAS_F32, float, true>(mg, st);
      LAUNCH_OK("d_pair");
    }
    {
      WgradBatch w; memset(&w, 0, sizeof(w));   // dW_p = d(e_p)^T zcat_p
      w.nbatch = 3; w.rows = B; w.M = D; w.N = 512; w.ldy = 3 * D; w.ldx = 512; w.ldo = 512; w.tf32 = TF32;
      bool any = false;
      for (int p = 0; p < 3; ++p) {
        w.dY[p] = decat + p * D; w.X[p] = zcat + (size_t)p * B * 512; w.out[p] = gr(ix.pair(p, 0));
        any = any || w.out[p] != nullptr;
      }
      if (any) { launch_wgrad_batched(w, ws); LAUNCH_OK("w_pair"); }
    }
    {
      ColsumArgs c; memset(&c, 0, sizeof(c));   // the three pair biases: column sums of d(ecat) [B, 768]
      c.segs = sb; c.src = decat; c.ld = 3 * D; c.col0 = 0; c.ncols = 3 * D; c.scale = 1.0f; c.colblock = D;
      bool any = false;
      for (int p = 0; p < 3; ++p) { c.out[p] = gr(ix.pair(p, 1)); any = any || c.out[p] != nullptr; }
      if (any) {
        colsum_kernel<float><<<dim3((3 * D + 255) / 256, (B + 127) / 128), 256, 0, ws>>>(c);
        LAUNCH_OK("b_pair");
      }
    }
  }
  // ---- pooled gradient through the encoder-final LayerNorm ----
  {
    LnBwdArgs a; memset(&a, 0, sizeof(a));
    a.q = Q; a.x = xin(L); a.stat = reinterpret_cast<const float*>(saved + P.s_statf); a.maskq = maskq;
    a.ld2 = 512; a.g_in = nullptr; a.g_out = g_a; a.gc_out = gc;
    for (int d = 0; d < NDIR; ++d) {
      a.dz[d] = d_routes + (size_t)route_of_dir(d) * B * D;
      a.dz2[d] = dzcat + (size_t)pair_of_dir(d) * B * 512 + half_of_dir(d) * D;
      a.cnt[d] = cnt + (size_t)dir_qmod(d) * B;
      a.gamma[d] = f(ix.enc_ln(d, 0));
      a.dgamma[d] = gr(ix.enc_ln(d, 0)); a.dbeta[d] = gr(ix.enc_ln(d, 1));
      a.dbias[d] = gr(ix.layer(d, L - 1, 7));
    }
    launch_k(ln_rows_bwd_kernel<CT, true>, dim3(P.MQ / 64), dim3(256), 0, st, a);
    LAUNCH_OK("lnf_bwd");
  }
  rc = main_to_side(E_GC0);   // fork: gc of the last layer is ready
  if (rc) return rc;
  const float* kmask[NDIR];
  for (int d = 0; d < NDIR; ++d) kmask[d] = mask[dir_kmod(d)];
  int maxTq = 0, maxTk = 0;
  for (int d = 0; d < NDIR; ++d) { maxTq = maxTq > Q.T[d] ? maxTq : Q.T[d]; maxTk = maxTk > P.kv.T[d] ? maxTk : P.kv.T[d]; }
  CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_dq_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_smem<AHG>()));
  CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_fused_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_fused_smem<AHG>()));
  CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_dkv_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_smem<AHG>()));

  auto q_problem = [&](const void* A, int lda, const void* Bw, int ldb, int nrows_b, int N, int K, int l) {
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = Q;
    for (int d = 0; d < NDIR; ++d) { g.a_row0[d] = Q.row0[d]; g.b_row0[d] = (l * 6 + d) * nrows_b; }
    g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = Bw; g.ldb = ldb;
    return g;
  };
  auto q_wgrad = [&](const void* dY, int ldy, int M, const void* X, int ldx, int N) {
    WgradProblem w; memset(&w, 0, sizeof(w));
    w.segs = Q;
    for (int d = 0; d < NDIR; ++d) w.x_row0[d] = Q.row0[d];
    w.dY = dY; w.ldy = ldy; w.X = X; w.ldx = ldx; w.M = M; w.N = N; w.ldo = N;
    return w;
  };

  float* g_cur = g_a;   // gradient wrt the layer output (masked)
  float* g_oth = g_b;
  for (int l = L - 1; l >= 0; --l) {
    {  // dW2 (reads gc, which the chain overwrites in LN1 backward)
      WgradProblem w = q_wgrad(gc, D, D, ff(l), FF, FF);
      for (int d = 0; d < NDIR; ++d) w.out[d] = gr(ix.layer(d, l, 6));
      rc = run_wgrad<CT>(P, w, P.MQ, P.MQ, ws, "w_fc2");
      if (rc) return rc;
      rc = side_record(S_DW2);
      if (rc) return rc;
    }
    if (have_dw1) { rc = main_wait(S_DW1); if (rc) return rc; }   // dW1 of the layer above still reads dF
    {  // dF = (G W2) .* relu'   (fc2 data gradient)
      GemmProblem g = q_problem(gc, D, packed + P.o_w2T, D, FF, FF, D, l);
      EpiSpec e;
      e.relu_src = ff(l); e.ld_relu = FF; e.out = dF; e.ldo = FF;
      e.bits = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(saved + P.s_bits + (size_t)l * P.l_bits)); e.ld_bits = FF / 32;
      rc = run_gemm<CT, G_RELU_BWD>(P, g, e, P.MQ, L * 6 * FF, st, "d_fc2");
      if (rc) return rc;
    }
    rc = main_to_side(E_DF);
    if (rc) return rc;
    {  // dH1 = dF W1, masked
      GemmProblem g = q_problem(dF, FF, packed + P.o_w1T, FF, D, D, FF, l);
      EpiSpec e;
      e.rowmask = maskq; e.out = dH; e.ldo = D;
      rc = run_gemm<CT, G_MASK>(P, g, e, P.MQ, L * 6 * D, st, "d_fc1");
      if (rc) return rc;
    }
    {  // dW1, db1
      WgradProblem w = q_wgrad(dF, FF, FF, h1(l), D, D);
      float* ob1[6];
      for (int d = 0; d < NDIR; ++d) { w.out[d] = gr(ix.layer(d, l, 4)); ob1[d] = gr(ix.layer(d, l, 5)); }
      const bool fused = fuse_colsum(P, w, ob1);
      rc = run_wgrad<CT>(P, w, P.MQ, P.MQ, ws, "w_fc1");
      if (rc) return rc;
      if (!fused) rc = run_colsum<CT>(Q, dF, FF, 0, FF, ob1, 1.0f, ws, "b_fc1");
      if (rc) return rc;
      rc = side_record(S_DW1);
      if (rc) return rc;
      have_dw1 = true;
    }
    if (have_dwo) { rc = main_wait(S_DWO); if (rc) return rc; }   // dWo of the layer above still reads the other gc buffer
    gci ^= 1; gc = gcbuf[gci];
    {  // LN1 backward: g_oth = (g_cur + dLN1) * mask ; d out_proj.bias = colsum(g_oth)
      LnBwdArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.dh = dH; a.x = x1(l); a.stat = stat1(l); a.maskq = maskq; a.g_in = g_cur; a.g_out = g_oth; a.gc_out = gc;
      for (int d = 0; d < NDIR; ++d) {
        a.gamma[d] = f(ix.layer(d, l, 10));
        a.dgamma[d] = gr(ix.layer(d, l, 10)); a.dbeta[d] = gr(ix.layer(d, l, 11));
        a.dbias[d] = gr(ix.layer(d, l, 3));
      }
      launch_k(ln_rows_bwd_kernel<CT, false>, dim3(P.MQ / 64), dim3(256), 0, st, a);
      LAUNCH_OK("ln1_bwd");
    }
    rc = main_to_side(E_GC1);
    if (rc) return rc;
    {  // dO = G1 Wo
      GemmProblem g = q_problem(gc, D, packed + P.o_woT, D, D, D, D, l);
      EpiSpec e;
      e.out = dO; e.ldo = D;
      rc = run_gemm<CT, G_MASK>(P, g, e, P.MQ, L * 6 * D, st, "d_out_proj");
      if (rc) return rc;
    }
    {  // dWo
      WgradProblem w = q_wgrad(gc, D, D, ob(l), D, D);
      for (int d = 0; d < NDIR; ++d) w.out[d] = gr(ix.layer(d, l, 2));
      rc = run_wgrad<CT>(P, w, P.MQ, P.MQ, ws, "w_out_proj");
      if (rc) return rc;
      rc = side_record(S_DWO);
      if (rc) return rc;
      have_dwo = true;
    }
    // out_proj / fc1 / fc2 weights+biases and LayerNorm-1 gradients of layer l (all six directions) are final
    // from here on: data-parallel callers may start reducing them while the rest of the backward runs
    if (layer_events && layer_events[l]) CUDA_OK(cudaEventRecord(reinterpret_cast<cudaEvent_t>(layer_events[l]), st));
    if (have_dwq) { rc = main_wait(S_DWQ); if (rc) return rc; }   // dWq of the layer above still reads dQ
    {  // attention backward
      AttnArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.kv = P.kv;
      for (int d = 0; d < NDIR; ++d) a.kmask[d] = kmask[d];
      a.qb = qb(l); a.kvbuf = kv; a.ldkv = ldkv; a.col0 = l * 2 * D; a.o = const_cast<CT*>(ob(l));
      a.ml = const_cast<float*>(ml(l)); a.d_o = dO; a.dq = dQ; a.dkv = dKV; a.dvec = dvec;
      bool pad_done = false;
      {
        ProfScope ps(PC_ATTN_BWD, st);
        if (mma_attention<CT>()) {
          if (maxTq <= amma::RC && maxTk <= amma::RC && !getenv("MMR_ATTN_BWD_SPLIT")) {
            // one chunk per sequence: fused dQ + dK/dV kernel (operands staged once, no O / D round trip)
            launch_k(amma::attn_bwd_fused_kernel<AHG>, dim3(amma::Cfg<AHG>::NHG, B, NDIR), dim3(amma::Cfg<AHG>::THREADS),
                     amma::bwd_fused_smem<AHG>(), st, a);
            pad_done = true;        // dQ's pad rows are cleared by the kernel
          } else {
          dim3 g1(amma::Cfg<AHG>::NHG * ((maxTq + amma::RC - 1) / amma::RC), B, NDIR);
          dim3 g2(amma::Cfg<AHG>::NHG * ((maxTk + amma::RC - 1) / amma::RC), B, NDIR);
          amma::attn_bwd_dq_kernel<AHG><<<g1, amma::Cfg<AHG>::THREADS, amma::bwd_smem<AHG>(), st>>>(a);
          amma::attn_bwd_dkv_kernel<AHG><<<g2, amma::Cfg<AHG>::THREADS, amma::bwd_smem<AHG>(), st>>>(a);
          }
        } else {
          dim3 g1((H * maxTq + ATT_THREADS - 1) / ATT_THREADS, B, NDIR);
          dim3 g2((H * maxTk + ATT_THREADS - 1) / ATT_THREADS, B, NDIR);
          attn_bwd_dq_kernel<CT><<<g1, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
          attn_bwd_dkv_kernel<CT><<<g2, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
        }
      }
      LAUNCH_OK("attn_bwd_dq");
      LAUNCH_OK("attn_bwd_dkv");
      if (!pad_done) {
        rc = zero_pad(Q, dQ, (size_t)D * sizeof(CT), st);
        if (rc) return rc;
      }
    }
    rc = main_to_side(E_DQ);
    if (rc) return rc;
    {  // dH0 = dQ Wq', masked
      GemmProblem g = q_problem(dQ, D, packed + P.o_wqT, D, D, D, D, l);
      EpiSpec e;
      e.rowmask = maskq; e.out = dH; e.ldo = D;
      rc = run_gemm<CT, G_MASK>(P, g, e, P.MQ, L * 6 * D, st, "d_q_proj");
      if (rc) return rc;
    }
    {  // dWq', dbq'
      WgradProblem w = q_wgrad(dQ, D, D, h0(l), D, D);
      float* obq[6];
      for (int d = 0; d < NDIR; ++d) { w.out[d] = dwq + ((size_t)l * 6 + d) * D * D; obq[d] = dbq + ((size_t)l * 6 + d) * D; }
      const bool fused = fuse_colsum(P, w, obq);
      rc = run_wgrad<CT>(P, w, P.MQ, P.MQ, ws, "w_q_proj");
      if (rc) return rc;
      if (!fused) rc = run_colsum<CT>(Q, dQ, D, 0, D, obq, 1.0f, ws, "b_q_proj");
      if (rc) return rc;
      rc = side_record(S_DWQ);
      if (rc) return rc;
      have_dwq = true;
    }
    rc = main_wait(S_DW2);   // LN0 backward overwrites the gc buffer dW2 of this layer reads
    if (rc) return rc;
    gci ^= 1; gc = gcbuf[gci];
    {  // LN0 backward: g_cur = (g_oth + dLN0) * mask ; d fc2.bias of the previous layer = colsum(g_cur)
      LnBwdArgs a; memset(&a, 0, sizeof(a));
      a.q = Q; a.dh = dH; a.x = xin(l); a.stat = stat0(l); a.maskq = maskq; a.g_in = g_oth; a.g_out = g_cur; a.gc_out = gc;
      for (int d = 0; d < NDIR; ++d) {
        a.gamma[d] = f(ix.layer(d, l, 8));
        a.dgamma[d] = gr(ix.layer(d, l, 8)); a.dbeta[d] = gr(ix.layer(d, l, 9));
        a.dbias[d] = l > 0 ? gr(ix.layer(d, l - 1, 7)) : nullptr;
      }
      launch_k(ln_rows_bwd_kernel<CT, false>, dim3(P.MQ / 64), dim3(256), 0, st, a);
      LAUNCH_OK("ln0_bwd");
    }
    rc = main_to_side(E_GC0);
    if (rc) return rc;
  }
  // ---- K/V stream ----
  rc = zero_pad(P.kv, dKV, (size_t)ldkv * sizeof(CT), st);
  if (rc) return rc;
  rc = main_to_side(E_DQ);   // dKV is complete
  if (rc) return rc;
  {
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = P.kv;
    for (int d = 0; d < NDIR; ++d) { g.a_row0[d] = P.kv.row0[d]; g.b_row0[d] = d * D; }
    g.N = D; g.K = ldkv; g.A = dKV; g.lda = ldkv; g.B = packed + P.o_wkvT; g.ldb = ldkv;
    EpiSpec e;
    e.out = dxh; e.ldo = D;
    rc = run_gemm<CT, G_F32>(P, g, e, P.MK, NDIR * D, st, "d_kv_proj");
    if (rc) return rc;
    WgradProblem w; memset(&w, 0, sizeof(w));
    w.segs = P.kv;
    float* obkv[6];
    for (int d = 0; d < NDIR; ++d) {
      w.x_row0[d] = P.mod.row0[dir_kmod(d)];
      w.out[d] = dwkv + (size_t)d * ldkv * D;
      obkv[d] = dbkv + (size_t)d * ldkv;
    }
    w.dY = dKV; w.ldy = ldkv; w.X = xh; w.ldx = D; w.M = ldkv; w.N = D; w.ldo = D;
    const bool fused = fuse_colsum(P, w, obkv);
    rc = run_wgrad<CT>(P, w, P.MK, P.MM, ws, "w_kv_proj");    // side stream: runs next to the dKV data-gradient GEMM
    if (rc) return rc;
    if (!fused) rc = run_colsum<CT>(P.kv, dKV, ldkv, 0, ldkv, obkv, 1.0f, ws, "b_kv_proj");
    if (rc) return rc;
  }
  // ---- unfold packed-weight gradients into in_proj_{weight,bias} and LN0 ----
  rc = side_record(S_JOIN);   // join: every weight gradient issued on the side stream is complete
  if (rc) return rc;
  rc = main_wait(S_JOIN);
  if (rc) return rc;
  {
    UnfoldJobs uj; memset(&uj, 0, sizeof(uj));
    uj.scaling = 1.0f / sqrtf((float)HD);
    for (int d = 0; d < NDIR; ++d)
      for (int l = 0; l < L; ++l) {
        UnfoldJob& j = uj.j[uj.n++];
        j.dwq = dwq + ((size_t)l * 6 + d) * D * D; j.dbq = dbq + ((size_t)l * 6 + d) * D;
        j.dwkv = dwkv + ((size_t)d * L + l) * 2 * D * D; j.dbkv = dbkv + ((size_t)d * L + l) * 2 * D;
        j.w_in = f(ix.layer(d, l, 0)); j.gamma0 = f(ix.layer(d, l, 8)); j.beta0 = f(ix.layer(d, l, 9));
        j.g_w_in = gr(ix.layer(d, l, 0)); j.g_b_in = gr(ix.layer(d, l, 1));
        j.g_gamma0 = gr(ix.layer(d, l, 8)); j.g_beta0 = gr(ix.layer(d, l, 9));
        if (uj.n == 24 || (d == NDIR - 1 && l == L - 1)) {
          unfold_kernel<<<dim3(3 * D / UNFOLD_ROWS, uj.n), 256, 0, st>>>(uj);
          LAUNCH_OK("unfold");
          uj.n = 0;
        }
      }
  }
  // ---- embedding backward ----
  {
    EmbedBwdArgs a; memset(&a, 0, sizeof(a));
    a.mod = P.mod; a.q = Q; a.kv = P.kv; a.xh = xh; a.rstd_e = reinterpret_cast<const float*>(saved + P.s_rstd_e);
    for (int m = 0; m < NMOD; ++m) a.tokrow[m] = reinterpret_cast<const int*>(saved + P.s_plan) + P.p_tokrow[m];
    a.g0 = g_cur; a.dxh = dxh; a.cnt = cnt; a.B = B;
    for (int m = 0; m < NMOD; ++m) {
      a.mask[m] = mask[m];
      a.dz_uni[m] = d_routes + (size_t)m * B * D;
      a.uni_g[m] = f(ix.uni_ln(m, 0));
      a.d_uni_g[m] = gr(ix.uni_ln(m, 0)); a.d_uni_b[m] = gr(ix.uni_ln(m, 1));
      a.dsrc[m] = (P.din[m] == D) ? dx[m] : dp + (size_t)P.mod.row0[m] * D;
    }
    embed_bwd_kernel<CT><<<P.MM / 64, 256, 0, st>>>(a);
    LAUNCH_OK("embed_bwd");
  }
  for (int m = 0; m < NMOD; ++m) {
    if (P.din[m] == D) continue;
    const float* dpm = dp + (size_t)P.mod.row0[m] * D;
    if (dx[m]) {
      rc = fp32_linear<true>(dpm, D, f(ix.proj(m)), P.din[m], nullptr, dx[m], P.din[m], P.mod.rows[m], P.din[m], D, st, "d_proj", TF32);
      if (rc) return rc;
    }
    rc = fp32_wgrad(dpm, D, x[m], P.din[m], gr(ix.proj(m)), P.din[m], P.mod.rows[m], D, P.din[m], st, "w_proj", TF32);
    if (rc) return rc;
  }
  return MMR_OK;
}

// ------------------------------------------------------- producer epilogue / training tail ---
template <class T>
static int sanitize_launch(bool bwd, const SanitizeArgs& a, cudaStream_t st) {
  const unsigned blocks = (unsigned)((a.rows + SAN_WARPS - 1) / SAN_WARPS);
  if (bwd) sanitize_bwd_kernel<T><<<blocks, SAN_WARPS * 32, 0, st>>>(a);
  else sanitize_fwd_kernel<T><<<blocks, SAN_WARPS * 32, 0, st>>>(a);
  LAUNCH_OK((bwd ? "sanitize_bwd" : "sanitize_fwd"));
  return MMR_OK;
}

static int sanitize_dispatch(bool bwd, const void* x, int in_dtype, float* y, const float* dy, float* dx,
                             int64_t rows, int Dw, int mode, float max_norm, unsigned long long* nonfinite,
                             void* stream) {
  if (rows < 0 || Dw <= 0 || Dw % 4 || Dw > SAN_MAX_D)
    return fail(MMR_ERR_INVALID_ARG, "sanitize: row width must be a positive multiple of 4 and <= 1024");
  if (mode != 0 && mode != 1) return fail(MMR_ERR_INVALID_ARG, "sanitize: mode must be 0 (clamp-norm) or 1 (nan_to_num)");
  if (!x || (bwd ? (!dy || !dx) : !y)) return fail(MMR_ERR_INVALID_ARG, "sanitize: null pointer");
  if (rows == 0) return MMR_OK;
  if ((rows + SAN_WARPS - 1) / SAN_WARPS > 0x7fffffffLL) return fail(MMR_ERR_INVALID_ARG, "sanitize: too many rows");
  SanitizeArgs a; memset(&a, 0, sizeof(a));
  a.x = x; a.y = y; a.dy = dy; a.dx = dx; a.rows = rows; a.D = Dw; a.mode = mode; a.max_norm = max_norm;
  a.nonfinite = nonfinite;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (in_dtype) {
    case MMR_DTYPE_F32: return sanitize_launch<float>(bwd, a, st);
    case MMR_DTYPE_BF16: return sanitize_launch<bf16>(bwd, a, st);
    case MMR_DTYPE_F16: return sanitize_launch<__half>(bwd, a, st);
  }
  return fail(MMR_ERR_INVALID_ARG, "sanitize: in_dtype must be MMR_DTYPE_F32 / BF16 / F16");
}

// Walks a host table of tensors in groups of OPT_NT and calls `launch(table, blocks)` for every group.
template <class F>
static int opt_for_tables(const mmr_opt_tensor* ts, int n, bool need_g, bool need_mv, bool need_ema, F&& launch) {
  if (n < 0 || (n > 0 && !ts)) return fail(MMR_ERR_INVALID_ARG, "optimizer table: null");
  OptTable t;
  int i = 0;
  while (i < n) {
    memset(&t, 0, sizeof(t));
    long long blocks = 0;
    while (i < n && t.nt < OPT_NT) {
      const mmr_opt_tensor& e = ts[i++];
      if (e.n < 0) return fail(MMR_ERR_INVALID_ARG, "optimizer table: negative size");
      if (e.n == 0) continue;
      if (!e.p || (need_g && !e.g) || (need_mv && (!e.m || !e.v)) || (need_ema && !e.ema))
        return fail(MMR_ERR_INVALID_ARG, "optimizer table: null tensor pointer");
      const int k = t.nt++;
      t.p[k] = e.p; t.g[k] = e.g; t.m[k] = e.m; t.v[k] = e.v; t.ema[k] = e.ema; t.n[k] = e.n;
      t.blk0[k] = (int)blocks;
      blocks += (e.n + OPT_CHUNK - 1) / OPT_CHUNK;
      if (blocks > 0x7fffffffLL) return fail(MMR_ERR_INVALID_ARG, "optimizer table: too many elements");
    }
    t.blk0[t.nt] = (int)blocks;
    if (t.nt == 0) continue;
    int rc = launch(t, (unsigned)blocks);
    if (rc) return rc;
  }
  return MMR_OK;
}


// ================================================================================ C ABI ===
extern "C" {

int mmr_sanitize_rows_fwd(const void* x, int in_dtype, float* y, int64_t rows, int Dw, int mode, float max_norm,
                          unsigned long long* nonfinite, void* stream) {
  return sanitize_dispatch(false, x, in_dtype, y, nullptr, nullptr, rows, Dw, mode, max_norm, nonfinite, stream);
}

int mmr_sanitize_rows_bwd(const void* x, int in_dtype, const float* dy, float* dx, int64_t rows, int Dw, int mode,
                          float max_norm, void* stream) {
  return sanitize_dispatch(true, x, in_dtype, nullptr, dy, dx, rows, Dw, mode, max_norm, nullptr, stream);
}

int mmr_route_mask_from_presence(const float* hasL, const float* hasN, const float* hasI, int B, int drop_bits,
                                 float* route_mask, void* stream) {
  if (B < 0 || !route_mask) return fail(MMR_ERR_INVALID_ARG, "mmr_route_mask_from_presence: bad arguments");
  if (drop_bits < 0 || drop_bits >= (1 << NR)) return fail(MMR_ERR_INVALID_ARG, "mmr_route_mask_from_presence: drop_bits has bits beyond the 10 routes");
  if (B == 0) return MMR_OK;
  route_mask_kernel<<<(B * NR + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hasL, hasN, hasI, B, drop_bits, route_mask);
  LAUNCH_OK("route_mask");
  return MMR_OK;
}

int mmr_grad_sqnorm(const mmr_opt_tensor* host_tensors, int n_tensors, mmr_opt_state* state, void* stream) {
  if (!state) return fail(MMR_ERR_INVALID_ARG, "mmr_grad_sqnorm: null state");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return opt_for_tables(host_tensors, n_tensors, true, false, false, [&](const OptTable& t, unsigned blocks) -> int {
    opt_sqnorm_kernel<<<blocks, OPT_THREADS, 0, st>>>(t, state);
    LAUNCH_OK("opt_sqnorm");
    return MMR_OK;
  });
}

int mmr_opt_prepare(const mmr_opt_hyper* hp, mmr_opt_state* state, void* stream) {
  if (!hp || !state) return fail(MMR_ERR_INVALID_ARG, "mmr_opt_prepare: null argument");
  if (!(hp->beta1 >= 0.0 && hp->beta1 < 1.0) || !(hp->beta2 >= 0.0 && hp->beta2 < 1.0))
    return fail(MMR_ERR_INVALID_ARG, "mmr_opt_prepare: betas must be in [0, 1)");
  opt_prepare_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(state, hp->beta1, hp->beta2, (float)hp->max_norm);
  LAUNCH_OK("opt_prepare");
  return MMR_OK;
}

int mmr_opt_apply(const mmr_opt_tensor* host_tensors, int n_tensors, const mmr_opt_hyper* hp,
                  const mmr_opt_state* state, void* stream) {
  if (!hp || !state) return fail(MMR_ERR_INVALID_ARG, "mmr_opt_apply: null argument");
  if (!(hp->lr >= 0.0) || !(hp->eps >= 0.0) || !(hp->weight_decay >= 0.0))
    return fail(MMR_ERR_INVALID_ARG, "mmr_opt_apply: lr, eps and weight_decay must be >= 0");
  const OptHyper h = opt_hyper(*hp);
  const double lr = hp->lr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return opt_for_tables(host_tensors, n_tensors, true, true, hp->has_ema != 0, [&](const OptTable& t, unsigned blocks) -> int {
    opt_step_kernel<<<blocks, OPT_THREADS, 0, st>>>(t, state, h, lr);
    LAUNCH_OK("opt_step");
    return MMR_OK;
  });
}

int mmr_ema_update(const mmr_opt_tensor* host_tensors, int n_tensors, double decay, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return opt_for_tables(host_tensors, n_tensors, false, false, true, [&](const OptTable& t, unsigned blocks) -> int {
    ema_update_kernel<<<blocks, OPT_THREADS, 0, st>>>(t, (float)decay, (float)(1.0 - decay));
    LAUNCH_OK("ema_update");
    return MMR_OK;
  });
}

// ------------------------------------------------------------------------------- loss tail ---
size_t mmr_loss_scratch_bytes(int B) {
  const size_t nblk = B > 0 ? (size_t)(B + LOSS_PB - 1) / LOSS_PB : 1;
  return nblk * LOSS_SLOTS * sizeof(double);
}

int mmr_loss_fwd_bwd(const mmr_loss_args* g, void* stream) {
  if (!g) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: null args");
  if (g->variant != MMR_VARIANT_MORT && g->variant != MMR_VARIANT_PHENO)
    return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: variant must be MMR_VARIANT_MORT or MMR_VARIANT_PHENO");
  if (g->B < 1) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: B must be >= 1");
  if (g->variant == MMR_VARIANT_MORT && g->K != 2)
    return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: the mortality head has 2 logits (death_logit_from_logits2 expects [B,2])");
  if (g->K < 1 || g->K > MMR_MAX_LABELS) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: K must be in [1, 32]");
  if (!g->logits || !g->y || !g->state || !g->scratch) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: null pointer");
  if ((reinterpret_cast<uintptr_t>(g->scratch) & 7) != 0) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: scratch must be 8-byte aligned");
  if (g->rc_dtype != MMR_DTYPE_F32 && g->rc_dtype != MMR_DTYPE_BF16)
    return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: rc_dtype must be MMR_DTYPE_F32 or MMR_DTYPE_BF16");
  if (g->variant == MMR_VARIANT_MORT && (g->rc_raw || g->rc_report || g->pos_weight))
    return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: rc_raw / rc_report / pos_weight belong to the PHENO variant");
  if (g->rc_report && !g->rc_raw) return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: rc_report needs rc_raw");
  if (!(g->label_smoothing >= 0.f && g->label_smoothing <= 1.f) || !(g->route_entropy_lambda >= 0.f) ||
      !(g->route_uniform_lambda >= 0.f) || !(g->atol > 0.f))
    return fail(MMR_ERR_INVALID_ARG, "mmr_loss_fwd_bwd: label_smoothing in [0,1], lambdas >= 0, atol > 0");
  LossArgs a;
  a.variant = g->variant; a.B = g->B; a.K = g->K;
  a.logits = g->logits; a.y = g->y; a.pos_weight = g->pos_weight; a.prim_acts = g->prim_acts;
  a.rc_raw = g->rc_raw; a.rc_bf16 = g->rc_dtype == MMR_DTYPE_BF16; a.route_mask = g->route_mask;
  a.label_smoothing = g->label_smoothing; a.lam_ent = g->route_entropy_lambda; a.lam_uni = g->route_uniform_lambda;
  a.atol = g->atol; a.dlogits = g->dlogits; a.rc_report = g->rc_report; a.state = g->state;
  a.scratch = reinterpret_cast<double*>(g->scratch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)((g->B + LOSS_PB - 1) / LOSS_PB);
  loss_stage1_kernel<<<blocks, LOSS_WARPS * 32, 0, st>>>(a);
  LAUNCH_OK("loss_stage1");
  if (g->variant == MMR_VARIANT_PHENO && g->rc_raw) {
    loss_stage2_kernel<<<blocks, LOSS_WARPS * 32, 0, st>>>(a);
    LAUNCH_OK("loss_stage2");
  }
  return MMR_OK;
}

int mmr_abi_struct_sizes(size_t* out, int n) {
  const size_t sz[] = {sizeof(mmr_fusion_dims), sizeof(mmr_routing_dims), sizeof(mmr_routing_params),
                       sizeof(mmr_routing_grads), sizeof(mmr_opt_tensor), sizeof(mmr_opt_hyper), sizeof(mmr_opt_state),
                       sizeof(mmr_loss_state), sizeof(mmr_loss_args), sizeof(mmr_proj_dims)};
  const int m = (int)(sizeof(sz) / sizeof(sz[0]));
  int i = 0;
  for (; out && i < n && i < m; ++i) out[i] = sz[i];
  return i;
}

// 100: route fusion + routing + tails; 101: + loss tail (mmr_loss_fwd_bwd); 102: + standalone projector, packed-weight
// forward, producer projections; 103: + mmr_capsule_routing_fwd_ex / _bwd_ex / mmr_routing_fwd_scratch_bytes (split routing path),
// mmr_attention_fwd / bwd (standalone attention core)
int mmr_version(void) { return 103; }

long long mmr_launch_count(void) { return g_launches.load(); }

int mmr_prof_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return MMR_OK;
}

// Synchronises the recorded events and returns the summed device time (ms) and launch-group count
// per class: 0 tcgen05 gemm, 1 tcgen05 wgrad, 2 attention fwd, 3 attention bwd, 4 SIMT gemm,
// 5 routing, 6 whole fusion fwd call, 7 whole fusion bwd call.
int mmr_prof_collect(double* ms_by_class, long long* n_by_class) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < PC_N; ++i) { if (ms_by_class) ms_by_class[i] = 0.0; if (n_by_class) n_by_class[i] = 0; }
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      if (ms_by_class) ms_by_class[r.cls] += ms;
      if (n_by_class) n_by_class[r.cls] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof_recs.clear();
  return MMR_OK;
}

const char* mmr_last_error_string(void) { return g_err.c_str(); }

int mmr_fusion_num_params(const mmr_fusion_dims* dims) {
  if (!dims || dims->layers < 1 || dims->layers > MMR_MAX_LAYERS) return -1;
  return ParamIndex{dims->layers}.count();
}

int mmr_fusion_sizes(const mmr_fusion_dims* dims, size_t* packed_bytes, size_t* saved_bytes, size_t* scratch_fwd_bytes,
                     size_t* scratch_bwd_bytes) {
  Plan P;
  const char* why = "";
  if (!build_plan(dims, &P, &why)) return fail(MMR_ERR_INVALID_ARG, why);
  if (packed_bytes) *packed_bytes = P.packed_bytes;
  if (saved_bytes) *saved_bytes = P.saved_bytes;
  if (scratch_fwd_bytes) *scratch_fwd_bytes = P.scratch_fwd_bytes;
  if (scratch_bwd_bytes) *scratch_bwd_bytes = P.scratch_bwd_bytes;
  return MMR_OK;
}

int mmr_fusion_pack_weights(const mmr_fusion_dims* dims, const void* const* host_params, void* packed, void* stream) {
  Plan P;
  const char* why = "";
  if (!build_plan(dims, &P, &why)) return fail(MMR_ERR_INVALID_ARG, why);
  if (!host_params || !packed) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.bf16) return pack_weights<bf16>(P, host_params, (uint8_t*)packed, st);
  return pack_weights<float>(P, host_params, (uint8_t*)packed, st);
}

static int route_fusion_fwd_impl(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                                 const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                                 const float* pos_table, void* packed, void* saved, void* scratch, float* routes_out,
                                 void* stream, bool do_pack) {
  Plan P;
  const char* why = "";
  if (!build_plan(dims, &P, &why)) return fail(MMR_ERR_INVALID_ARG, why);
  if (!host_params || !x_l || !x_n || !x_i || !pos_table || !packed || !saved || !scratch || !routes_out)
    return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  const float* x[3] = {x_l, x_n, x_i};
  const float* mask[3] = {mL, mN, mI};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope ps(PC_FUSION_FWD, st);
  if (P.bf16)
    return fusion_fwd<bf16>(P, host_params, x, mask, pos_table, (uint8_t*)packed, (uint8_t*)saved, (uint8_t*)scratch,
                            routes_out, st, do_pack);
  return fusion_fwd<float>(P, host_params, x, mask, pos_table, (uint8_t*)packed, (uint8_t*)saved, (uint8_t*)scratch,
                           routes_out, st, do_pack);
}

int mmr_route_fusion_fwd(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                         const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                         const float* pos_table, void* packed, void* saved, void* scratch, float* routes_out,
                         void* stream) {
  return route_fusion_fwd_impl(dims, host_params, x_l, x_n, x_i, mL, mN, mI, pos_table, packed, saved, scratch, routes_out,
                               stream, true);
}

int mmr_route_fusion_fwd_packed(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                                const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                                const float* pos_table, const void* packed, void* saved, void* scratch, float* routes_out,
                                void* stream) {
  return route_fusion_fwd_impl(dims, host_params, x_l, x_n, x_i, mL, mN, mI, pos_table, const_cast<void*>(packed), saved,
                               scratch, routes_out, stream, false);
}

int mmr_route_fusion_bwd_ex(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                            const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                            const void* packed, const void* saved, void* scratch, const float* d_routes,
                            void* const* host_param_grads, float* dx_l, float* dx_n, float* dx_i, void* stream,
                            void* const* host_layer_events, void* side_stream, void* const* host_sync_events) {
  Plan P;
  const char* why = "";
  if (!build_plan(dims, &P, &why)) return fail(MMR_ERR_INVALID_ARG, why);
  if (!host_params || !x_l || !x_n || !x_i || !packed || !saved || !scratch || !d_routes || !host_param_grads)
    return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (side_stream && !host_sync_events) return fail(MMR_ERR_INVALID_ARG, "side_stream needs MMR_BWD_SYNC_EVENTS events");
  if (side_stream && side_stream == stream) return fail(MMR_ERR_INVALID_ARG, "side_stream must differ from stream");
  if (side_stream)
    for (int i = 0; i < MMR_BWD_SYNC_EVENTS; ++i)
      if (!host_sync_events[i]) return fail(MMR_ERR_INVALID_ARG, "null entry in host_sync_events");
  const float* x[3] = {x_l, x_n, x_i};
  const float* mask[3] = {mL, mN, mI};
  float* dx[3] = {dx_l, dx_n, dx_i};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaStream_t side = reinterpret_cast<cudaStream_t>(side_stream);
  ProfScope ps(PC_FUSION_BWD, st);
  if (P.bf16)
    return fusion_bwd<bf16>(P, host_params, x, mask, (const uint8_t*)packed, (const uint8_t*)saved, (uint8_t*)scratch,
                            d_routes, host_param_grads, dx, st, host_layer_events, side, host_sync_events);
  return fusion_bwd<float>(P, host_params, x, mask, (const uint8_t*)packed, (const uint8_t*)saved, (uint8_t*)scratch,
                           d_routes, host_param_grads, dx, st, host_layer_events, side, host_sync_events);
}

int mmr_route_fusion_bwd_events(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                                const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                                const void* packed, const void* saved, void* scratch, const float* d_routes,
                                void* const* host_param_grads, float* dx_l, float* dx_n, float* dx_i, void* stream,
                                void* const* host_layer_events) {
  return mmr_route_fusion_bwd_ex(dims, host_params, x_l, x_n, x_i, mL, mN, mI, packed, saved, scratch, d_routes,
                                 host_param_grads, dx_l, dx_n, dx_i, stream, host_layer_events, nullptr, nullptr);
}

int mmr_route_fusion_bwd(const mmr_fusion_dims* dims, const void* const* host_params, const float* x_l,
                         const float* x_n, const float* x_i, const float* mL, const float* mN, const float* mI,
                         const void* packed, const void* saved, void* scratch, const float* d_routes,
                         void* const* host_param_grads, float* dx_l, float* dx_n, float* dx_i, void* stream) {
  return mmr_route_fusion_bwd_ex(dims, host_params, x_l, x_n, x_i, mL, mN, mI, packed, saved, scratch, d_routes,
                                 host_param_grads, dx_l, dx_n, dx_i, stream, nullptr, nullptr, nullptr);
}

// ------------------------------------------------------------------------------- routing ---
static int check_routing(const mmr_routing_dims* d) {
  if (!d) return fail(MMR_ERR_INVALID_ARG, "dims is NULL");
  if (d->B <= 0) return fail(MMR_ERR_INVALID_ARG, "B must be positive");
  if (d->K < 1 || d->K > MMR_MAX_LABELS) return fail(MMR_ERR_UNSUPPORTED, "K must be in [1,32]");
  if (d->num_routing < 1 || d->num_routing > RT_MAXIT) return fail(MMR_ERR_UNSUPPORTED, "num_routing must be in [1,4]");
  if (d->variant != MMR_VARIANT_MORT && d->variant != MMR_VARIANT_PHENO) return fail(MMR_ERR_INVALID_ARG, "unknown variant");
  if (!(d->act_temperature > 0.f)) return fail(MMR_ERR_INVALID_ARG, "act_temperature must be > 0");
  if (d->vote_dtype != MMR_DTYPE_F32 && d->vote_dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_INVALID_ARG, "unknown vote_dtype");
  return MMR_OK;
}

// backward scratch: du | dpc | posem | dG | (split path) pose | zl | G | votes | dposeA;  forward scratch: pose | zl | G | votes
static size_t rs_fwd_scratch(size_t B, size_t K) {
  return align256(B * 320 * 4) + align256(B * 10 * 4) + align256(K * 32 * 4) + align256(B * 10 * K * 64 * 2) +
         align256(B * (RS_NIT - 1) * 320 * 4);
}
size_t mmr_routing_scratch_bytes(const mmr_routing_dims* d) {
  if (!d || d->B <= 0 || d->K < 1) return 0;
  const size_t B = d->B, K = d->K;
  return align256(B * 10 * K * 64 * 4) + align256(B * 330 * 4) + align256(B * 320 * 4) + align256(K * 32 * 4) +
         rs_fwd_scratch(B, K) + align256(B * 320 * 4) + align256((size_t)RS_GCOPIES * (K + 1) * 32 * 4) + 256;
}
size_t mmr_routing_fwd_scratch_bytes(const mmr_routing_dims* d) {
  if (!d || d->B <= 0 || d->K < 1) return 0;
  return rs_fwd_scratch(d->B, d->K) + 256;
}

}  // extern "C"

// Tile size (patients per CTA) and vote storage type.  The largest tile that fits in shared memory is
// used as long as it still leaves >= 100 tiles (so the SMs stay busy); small batches use small tiles.
struct RtLaunch { int PB; bool bf16; size_t smem; int grid; };
static RtLaunch routing_launch_cfg(const mmr_routing_dims* d, bool bwd) {
  RtLaunch L;
  L.bf16 = d->vote_dtype == MMR_DTYPE_BF16;
  const size_t ut = L.bf16 ? 2 : 4;
  const int cand[4] = {8, 4, 2, 1};
  L.PB = 1;
  const char* force = getenv("MMR_RT_PB");   // tuning override: largest tile size to consider
  const int pb_max = force ? atoi(force) : 8;
  for (int i = 0; i < 4; ++i) {
    const int pb = cand[i];
    if (pb > pb_max) continue;
    const size_t smem = rt_smem_bytes(d->K, d->num_routing, bwd, pb, ut, d->from_poses != 0);
    if (smem > (size_t)227 * 1024) continue;
    if (pb > 1 && (d->B + pb - 1) / pb < 100) continue;
    L.PB = pb;
    break;
  }
  L.smem = rt_smem_bytes(d->K, d->num_routing, bwd, L.PB, ut, d->from_poses != 0);
  int per_sm = (int)((size_t)(227 * 1024) / (L.smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  const int ntiles = (d->B + L.PB - 1) / L.PB;
  L.grid = ntiles < 148 * per_sm ? ntiles : 148 * per_sm;
  return L;
}

template <int PB, class UT>
static cudaError_t launch_routing(const RoutingArgs& a, const RtLaunch& L, bool bwd, cudaStream_t st) {
  cudaError_t e;
  if (bwd) {
    e = cudaFuncSetAttribute(routing_bwd_kernel<PB, UT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    routing_bwd_kernel<PB, UT><<<L.grid, RT_THREADS, L.smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(routing_fwd_kernel<PB, UT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.smem);
    if (e != cudaSuccess) return e;
    routing_fwd_kernel<PB, UT><<<L.grid, RT_THREADS, L.smem, st>>>(a);
  }
  return cudaGetLastError();
}
static cudaError_t dispatch_routing(const RoutingArgs& a, const RtLaunch& L, bool bwd, cudaStream_t st) {
#define MMR_RT_CASE(pb)                                                          \
  case pb:                                                                       \
    return L.bf16 ? launch_routing<pb, __half>(a, L, bwd, st) : launch_routing<pb, float>(a, L, bwd, st);
  switch (L.PB) {
    MMR_RT_CASE(1) MMR_RT_CASE(2) MMR_RT_CASE(4) MMR_RT_CASE(8)
  }
#undef MMR_RT_CASE
  return cudaErrorInvalidValue;
}

// ---- split path (routing_split.cuh): projector / votes GEMMs + one patient per 1-4 warps for the agreement iterations ----
static bool rs_enabled(const mmr_routing_dims* d, const mmr_routing_params* p, const void* scratch) {
  const char* e = getenv("MMR_RT_SPLIT");      // read per call: the tests switch paths inside one process
  const bool on = !(e && atoi(e) == 0);
  return on && scratch && d->vote_dtype == MMR_DTYPE_BF16 && p->caps_wt_f16 && p->caps_w_f16 &&
         (d->from_poses || p->proj_w_f16) && d->num_routing <= RS_NIT;
}
static uint8_t* rs_carve_fwd(uint8_t* s, size_t B, size_t K, RsScratch* o) {
  o->pose = reinterpret_cast<float*>(s); s += align256(B * 320 * 4);
  o->zl = reinterpret_cast<float*>(s); s += align256(B * 10 * 4);
  o->G = reinterpret_cast<float*>(s); s += align256(K * 32 * 4);
  o->votes = reinterpret_cast<__half*>(s); s += align256(B * 10 * K * 64 * 2);
  o->qs = reinterpret_cast<float*>(s); s += align256(B * (RS_NIT - 1) * 320 * 4);
  o->dposeA = nullptr; o->dGc = nullptr;
  return s;
}
template <int KP>
static cudaError_t rs_launch_iterate(const RoutingArgs& a, const RsScratch& sc, bool bwd, cudaStream_t st) {
  using C = RsCfg<KP>;
  const size_t per = (rs_patient_smem(a.d.K, KP, C::NW) + 15) / 16 * 16, smem = per * C::PPC;
  int per_sm = (int)((size_t)(227 * 1024) / (smem + 1024));
  const int cap = bwd ? RS_BWD_MINB : 4;
  per_sm = per_sm < 1 ? 1 : (per_sm > cap ? cap : per_sm);
  const int units = (a.d.B + C::PPC - 1) / C::PPC;
  const int grid = units < 148 * per_sm ? units : 148 * per_sm;
  cudaError_t e;
  if (bwd) {
    e = cudaFuncSetAttribute(rs_iterate_bwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rs_iterate_bwd_kernel<KP><<<grid, 128, smem, st>>>(a, sc);
  } else {
    e = cudaFuncSetAttribute(rs_iterate_fwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    rs_iterate_fwd_kernel<KP><<<grid, 128, smem, st>>>(a, sc);
  }
  return cudaGetLastError();
}
static cudaError_t rs_dispatch_iterate(const RoutingArgs& a, const RsScratch& sc, bool bwd, cudaStream_t st) {
  const int K = a.d.K;
  if (K <= 2) return rs_launch_iterate<2>(a, sc, bwd, st);
  if (K <= 4) return rs_launch_iterate<4>(a, sc, bwd, st);
  if (K <= 8) return rs_launch_iterate<8>(a, sc, bwd, st);
  if (K <= 16) return rs_launch_iterate<16>(a, sc, bwd, st);
  return rs_launch_iterate<32>(a, sc, bwd, st);
}
// projector (+ G) and votes; returns the number of launches through *n
static cudaError_t rs_launch_front(const RoutingArgs& a, const RsScratch& sc, cudaStream_t st, int* n) {
  const int tiles = (a.d.B + 15) / 16, NU = a.d.K * 4;
  rs_project_kernel<<<dim3(a.d.from_poses ? 1 : tiles, 11), 128, 0, st>>>(a, sc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  int S = (2 * 148 + tiles - 1) / tiles;
  if (S > NU / 8) S = NU / 8;
  if (S < 1) S = 1;
  rs_votes_kernel<<<dim3(tiles, S), 256, 0, st>>>(a, sc);
  *n += 2;
  return cudaGetLastError();
}

extern "C" {

int mmr_routing_pack_weights(const mmr_routing_params* params, int K, void* caps_wt_f16, void* caps_w_f16,
                             void* proj_w_f16, void* stream) {
  if (!params || !params->caps_w || !caps_wt_f16 || !caps_w_f16) return fail(MMR_ERR_INVALID_ARG, "mmr_routing_pack_weights: null argument");
  if (K < 1 || K > MMR_MAX_LABELS) return fail(MMR_ERR_INVALID_ARG, "mmr_routing_pack_weights: K out of range");
  const bool proj = proj_w_f16 != nullptr && params->proj_w[0] != nullptr;
  if (proj)
    for (int r = 0; r < NR; ++r)
      if (!params->proj_w[r]) return fail(MMR_ERR_INVALID_ARG, "mmr_routing_pack_weights: missing projector weight");
  const long long n = 10LL * K * 64 + 10LL * 32 * K * 64 / 8 + (proj ? 10LL * 40 * 32 : 0);
  routing_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      *params, K, reinterpret_cast<__half*>(caps_wt_f16), reinterpret_cast<__half*>(caps_w_f16), proj ? reinterpret_cast<__half*>(proj_w_f16) : nullptr);
  LAUNCH_OK("routing_pack");
  return MMR_OK;
}

int mmr_capsule_routing_fwd(const mmr_routing_dims* dims, const mmr_routing_params* params, const float* route_embs,
                            const float* poses_in, const float* acts_in, const float* acts_override,
                            const float* route_mask, float* logits, float* alpha, float* R, float* poses_out,
                            float* acts_out, void* stream) {
  return mmr_capsule_routing_fwd_ex(dims, params, route_embs, poses_in, acts_in, acts_override, route_mask, logits, alpha, R,
                                    poses_out, acts_out, nullptr, stream);
}

int mmr_capsule_routing_fwd_ex(const mmr_routing_dims* dims, const mmr_routing_params* params, const float* route_embs,
                               const float* poses_in, const float* acts_in, const float* acts_override,
                               const float* route_mask, float* logits, float* alpha, float* R, float* poses_out,
                               float* acts_out, void* scratch, void* stream) {
  int rc = check_routing(dims);
  if (rc) return rc;
  if (!params || !logits || !alpha) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (dims->from_poses ? (!poses_in || !acts_in) : !route_embs) return fail(MMR_ERR_INVALID_ARG, "missing routing input");
  RoutingArgs a; memset(&a, 0, sizeof(a));
  a.d = *dims; a.p = *params;
  a.route_embs = route_embs; a.poses_in = poses_in; a.acts_in = acts_in; a.acts_override = acts_override;
  a.route_mask = route_mask; a.logits = logits; a.alpha = alpha; a.R = R; a.poses_out = poses_out; a.acts_out = acts_out;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ProfScope ps(PC_ROUTING, st);
  if (rs_enabled(dims, params, scratch)) {
    RsScratch sc;
    rs_carve_fwd(reinterpret_cast<uint8_t*>(scratch), dims->B, dims->K, &sc);
    int n = 0;
    CUDA_OK(rs_launch_front(a, sc, st, &n));
    CUDA_OK(rs_dispatch_iterate(a, sc, false, st));
    g_launches.fetch_add(n + 1, std::memory_order_relaxed);
    return MMR_OK;
  }
  const RtLaunch L = routing_launch_cfg(dims, false);
  CUDA_OK(dispatch_routing(a, L, false, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMR_OK;
}

__global__ void proj_bias_grad_kernel(const float* dpc, int B, mmr_routing_grads g) {
  __shared__ float red[8][33];
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  for (int b = warp; b < B; b += 8) {
    const float* p = dpc + (size_t)b * 330 + r * 33;
    a0 += p[lane];
    if (lane == 0) a1 += p[32];
  }
  red[warp][lane] = a0;
  if (lane == 0) red[warp][32] = a1;
  __syncthreads();
  if (threadIdx.x < 33 && g.proj_b[r]) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    g.proj_b[r][threadIdx.x] += s;
  }
}

int mmr_capsule_routing_bwd(const mmr_routing_dims* dims, const mmr_routing_params* params, const float* route_embs,
                            const float* poses_in, const float* acts_in, const float* acts_override,
                            const float* route_mask, const float* d_logits, const float* d_R, void* scratch,
                            const mmr_routing_grads* grads, float* d_route_embs, float* d_poses, float* d_acts,
                            void* stream) {
  return mmr_capsule_routing_bwd_ex(dims, params, route_embs, poses_in, acts_in, acts_override, route_mask, d_logits, d_R,
                                    scratch, grads, d_route_embs, d_poses, d_acts, nullptr, stream);
}

int mmr_capsule_routing_bwd_ex(const mmr_routing_dims* dims, const mmr_routing_params* params, const float* route_embs,
                               const float* poses_in, const float* acts_in, const float* acts_override,
                               const float* route_mask, const float* d_logits, const float* d_R, void* scratch,
                               const mmr_routing_grads* grads, float* d_route_embs, float* d_poses, float* d_acts,
                               const void* fwd_scratch, void* stream) {
  int rc = check_routing(dims);
  if (rc) return rc;
  if (!params || !d_logits || !scratch || !grads) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (dims->from_poses ? (!poses_in || !acts_in) : !route_embs) return fail(MMR_ERR_INVALID_ARG, "missing routing input");
  const size_t B = dims->B, K = dims->K, KD = K * 64;
  uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
  float* du = reinterpret_cast<float*>(s); s += align256(B * 10 * KD * 4);
  float* dpc = reinterpret_cast<float*>(s); s += align256(B * 330 * 4);
  float* posem = reinterpret_cast<float*>(s); s += align256(B * 320 * 4);
  float* dG = reinterpret_cast<float*>(s); s += align256(K * 32 * 4);
  RsScratch sc;
  s = rs_carve_fwd(s, B, K, &sc);
  sc.dposeA = reinterpret_cast<float*>(s); s += align256(B * 320 * 4);
  sc.dGc = reinterpret_cast<float*>(s);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CUDA_OK(cudaMemsetAsync(dG, 0, K * 32 * 4, st));
  RoutingArgs a; memset(&a, 0, sizeof(a));
  a.d = *dims; a.p = *params;
  a.route_embs = route_embs; a.poses_in = poses_in; a.acts_in = acts_in; a.acts_override = acts_override;
  a.route_mask = route_mask; a.d_logits = d_logits; a.d_R = d_R;
  a.d_route_embs = d_route_embs; a.d_poses = d_poses; a.d_acts = d_acts;
  a.du = du; a.dpc = dpc; a.dG = dG; a.dbias = grads->bias; a.poses_m = posem;
  ProfScope ps(PC_ROUTING, st);
  bool split = false;
  if (rs_enabled(dims, params, scratch)) {
    int n = 0;
    CUDA_OK(cudaMemsetAsync(sc.dGc, 0, (size_t)RS_GCOPIES * (K + 1) * 32 * 4, st));
    if (fwd_scratch) {     // projector outputs, head matrix and votes of the forward call: nothing to recompute
      float* dposeA = sc.dposeA; float* dGc = sc.dGc;
      rs_carve_fwd(reinterpret_cast<uint8_t*>(const_cast<void*>(fwd_scratch)), B, K, &sc);
      sc.dposeA = dposeA; sc.dGc = dGc;
    } else {
      sc.qs = nullptr;       // recompute path: the agreement iterations run again inside the backward kernel
      CUDA_OK(rs_launch_front(a, sc, st, &n));
    }
    CUDA_OK(rs_dispatch_iterate(a, sc, true, st));
    rs_fold_copies_kernel<<<((int)(K + 1) * 32 + 255) / 256, 256, 0, st>>>(sc.dGc, (int)K, dG, grads->bias);
    CUDA_OK(cudaGetLastError());
    for (int r = 0; r < 10; ++r) a.d_proj_b[r] = dims->from_poses ? nullptr : grads->proj_b[r];
    rs_dpose_kernel<<<dim3((dims->B + 15) / 16, 10), 256, 0, st>>>(a, sc);
    CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(n + 3, std::memory_order_relaxed);
    split = true;
  } else {
    const RtLaunch L = routing_launch_cfg(dims, true);
    CUDA_OK(dispatch_routing(a, L, true, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  routing_head_grads_kernel<<<(MC * PC + (int)K * MC + 255) / 256, 256, 0, st>>>(dG, params->pose_to_mc, params->embedding, dims->K, grads->pose_to_mc,
                                               grads->embedding);
  LAUNCH_OK("routing_head_grads");
  if (grads->caps_w) {   // d w[r] = pose_masked[:, r, :]^T du[:, r, :]
    WgradBatch w; memset(&w, 0, sizeof(w));
    w.nbatch = 10; w.rows = dims->B; w.M = 32; w.N = (int)KD; w.ldy = 320; w.ldx = (int)(10 * KD); w.ldo = (int)KD;
    w.tf32 = dims->vote_dtype == MMR_DTYPE_BF16 && !(getenv("MMR_TF32") && atoi(getenv("MMR_TF32")) == 0);
    for (int r = 0; r < 10; ++r) { w.dY[r] = posem + r * 32; w.X[r] = du + r * KD; w.out[r] = grads->caps_w + (size_t)r * 32 * KD; }
    launch_wgrad_batched(w, st);
    LAUNCH_OK("w_caps");
  }
  if (!dims->from_poses) {
    WgradBatch w; memset(&w, 0, sizeof(w));
    w.nbatch = 10; w.rows = dims->B; w.M = 33; w.N = 256; w.ldy = 330; w.ldx = (int)dims->emb_batch_stride; w.ldo = 256;
    w.tf32 = dims->vote_dtype == MMR_DTYPE_BF16 && !(getenv("MMR_TF32") && atoi(getenv("MMR_TF32")) == 0);
    bool any = false;
    for (int r = 0; r < 10; ++r) {
      w.dY[r] = dpc + r * 33; w.X[r] = route_embs + (size_t)r * dims->emb_route_stride; w.out[r] = grads->proj_w[r];
      any = any || grads->proj_w[r] != nullptr;
    }
    if (any) { launch_wgrad_batched(w, st); LAUNCH_OK("w_proj"); }
    if (!split) {      // the split path accumulates the bias gradient inside rs_dpose_kernel
      proj_bias_grad_kernel<<<10, 256, 0, st>>>(dpc, dims->B, *grads);
      LAUNCH_OK("b_proj");
    }
  }
  return MMR_OK;
}

// ------------------------------------------------------------------- standalone projector ---
static int projector_args(ProjectorArgs* a, const mmr_routing_params* params, const float* route_embs, int64_t rs, int64_t bs,
                          int B) {
  if (!params || !route_embs) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (B <= 0 || bs < 256 || rs < 0) return fail(MMR_ERR_INVALID_ARG, "bad projector batch / strides");
  memset(a, 0, sizeof(*a));
  for (int r = 0; r < MMR_ROUTES; ++r) {
    if (!params->proj_w[r] || !params->proj_b[r]) return fail(MMR_ERR_INVALID_ARG, "null projector weight");
    a->w[r] = params->proj_w[r]; a->b[r] = params->proj_b[r];
  }
  a->embs = route_embs; a->rs = rs; a->bs = bs; a->B = B;
  return MMR_OK;
}

int mmr_projector_fwd(const mmr_routing_params* params, const float* route_embs, int64_t emb_route_stride,
                      int64_t emb_batch_stride, int B, float* poses, float* acts, void* stream) {
  ProjectorArgs a;
  int rc = projector_args(&a, params, route_embs, emb_route_stride, emb_batch_stride, B);
  if (rc) return rc;
  if (!poses || !acts) return fail(MMR_ERR_INVALID_ARG, "null output");
  a.poses = poses; a.acts = acts;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  projector_fwd_kernel<<<dim3((B + PJ_PB - 1) / PJ_PB, MMR_ROUTES), 256, 0, st>>>(a);
  LAUNCH_OK("projector_fwd");
  return MMR_OK;
}

int mmr_projector_bwd(const mmr_routing_params* params, const float* route_embs, int64_t emb_route_stride,
                      int64_t emb_batch_stride, int B, const float* d_poses, const float* d_acts, void* scratch,
                      const mmr_routing_grads* grads, float* d_route_embs, void* stream) {
  ProjectorArgs a;
  int rc = projector_args(&a, params, route_embs, emb_route_stride, emb_batch_stride, B);
  if (rc) return rc;
  if (!scratch || !grads) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  float* dpc = reinterpret_cast<float*>(scratch);     // [B,10,33]
  a.d_poses = d_poses; a.d_acts = d_acts; a.dpc = dpc; a.d_embs = d_route_embs;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  projector_bwd_kernel<<<dim3((B + PJ_PB - 1) / PJ_PB, MMR_ROUTES), 256, 0, st>>>(a);
  LAUNCH_OK("projector_bwd");
  WgradBatch w; memset(&w, 0, sizeof(w));
  w.nbatch = 10; w.rows = B; w.M = 33; w.N = 256; w.ldy = 330; w.ldx = (int)emb_batch_stride; w.ldo = 256;
  bool any = false;
  for (int r = 0; r < 10; ++r) {
    w.dY[r] = dpc + r * 33; w.X[r] = route_embs + (size_t)r * emb_route_stride; w.out[r] = grads->proj_w[r];
    any = any || grads->proj_w[r] != nullptr;
  }
  if (any) { launch_wgrad_batched(w, st); LAUNCH_OK("w_proj"); }
  proj_bias_grad_kernel<<<10, 256, 0, st>>>(dpc, B, *grads);
  LAUNCH_OK("b_proj");
  return MMR_OK;
}

// ---------------------------------------------------------------- standalone attention core ---
// The path's per-patient attention kernels for ONE (query, key) pair of streams: the consumers outside MULTModel
// (the Partial/ attention-fusion variants, multimodalrouting_b200/partial_fusion.py) reach them through these two calls.
}  // extern "C"

template <class CT>
static int attention_core(bool bwd, int B, int Tq, int Tk, const void* q, const void* kv, const float* kmask, void* o, float* ml,
                          const void* d_o, void* dq, void* dkv, float* dvec, cudaStream_t st) {
  AttnArgs a; memset(&a, 0, sizeof(a));
  a.q = single_seg(B * Tq, Tq); a.kv = single_seg(B * Tk, Tk);
  a.kmask[0] = kmask;
  a.qb = q; a.kvbuf = kv; a.ldkv = 2 * D; a.col0 = 0; a.o = o; a.ml = ml;
  a.d_o = d_o; a.dq = dq; a.dkv = dkv; a.dvec = dvec;
  const bool mma = std::is_same<CT, bf16>::value;
  if (!bwd) {
    ProfScope ps(PC_ATTN_FWD, st);
    if (mma) {
      CUDA_OK(cudaFuncSetAttribute(amma::attn_fwd_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::fwd_smem<AHG>()));
      CUDA_OK(cudaFuncSetAttribute(amma::attn_fwd_single_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::fwd_smem<AHG>()));
      dim3 grid(amma::Cfg<AHG>::NHG * ((Tq + amma::RC - 1) / amma::RC), B, 1);
      if (Tk <= amma::RC) amma::attn_fwd_single_kernel<AHG><<<grid, amma::Cfg<AHG>::THREADS, amma::fwd_smem<AHG>(), st>>>(a);
      else launch_k(amma::attn_fwd_kernel<AHG>, grid, dim3(amma::Cfg<AHG>::THREADS), amma::fwd_smem<AHG>(), st, a);
    } else {
      CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
      dim3 grid((H * Tq + ATT_THREADS - 1) / ATT_THREADS, B, 1);
      attn_fwd_kernel<CT><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
    }
    LAUNCH_OK("attention_fwd");
    return MMR_OK;
  }
  ProfScope ps(PC_ATTN_BWD, st);
  if (mma) {
    CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_dq_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_smem<AHG>()));
    CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_fused_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_fused_smem<AHG>()));
    CUDA_OK(cudaFuncSetAttribute(amma::attn_bwd_dkv_kernel<AHG>, cudaFuncAttributeMaxDynamicSharedMemorySize, amma::bwd_smem<AHG>()));
    if (Tq <= amma::RC && Tk <= amma::RC) {
      launch_k(amma::attn_bwd_fused_kernel<AHG>, dim3(amma::Cfg<AHG>::NHG, B, 1), dim3(amma::Cfg<AHG>::THREADS),
               amma::bwd_fused_smem<AHG>(), st, a);
    } else {
      dim3 g1(amma::Cfg<AHG>::NHG * ((Tq + amma::RC - 1) / amma::RC), B, 1);
      dim3 g2(amma::Cfg<AHG>::NHG * ((Tk + amma::RC - 1) / amma::RC), B, 1);
      amma::attn_bwd_dq_kernel<AHG><<<g1, amma::Cfg<AHG>::THREADS, amma::bwd_smem<AHG>(), st>>>(a);
      amma::attn_bwd_dkv_kernel<AHG><<<g2, amma::Cfg<AHG>::THREADS, amma::bwd_smem<AHG>(), st>>>(a);
    }
  } else {
    CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    dim3 g1((H * Tq + ATT_THREADS - 1) / ATT_THREADS, B, 1);
    dim3 g2((H * Tk + ATT_THREADS - 1) / ATT_THREADS, B, 1);
    attn_bwd_dq_kernel<CT><<<g1, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
    attn_bwd_dkv_kernel<CT><<<g2, ATT_THREADS, ATT_SMEM_BYTES, st>>>(a);
  }
  LAUNCH_OK("attention_bwd");
  return MMR_OK;
}

extern "C" {

static int check_attention(int dtype, int B, int Tq, int Tk, const void* q, const void* kv) {
  if (dtype != MMR_DTYPE_F32 && dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_INVALID_ARG, "attention: dtype must be F32 or BF16");
  if (B <= 0 || Tq <= 0 || Tk <= 0) return fail(MMR_ERR_INVALID_ARG, "attention: B, Tq, Tk must be positive");
  if (Tq > 4096 || Tk > 4096) return fail(MMR_ERR_UNSUPPORTED, "attention: sequences longer than 4096 tokens");
  if (!q || !kv) return fail(MMR_ERR_INVALID_ARG, "attention: null operand");
  return MMR_OK;
}

int mmr_attention_fwd(int dtype, int B, int Tq, int Tk, const void* q, const void* kv, const float* kmask, void* o, float* ml,
                      void* stream) {
  int rc = check_attention(dtype, B, Tq, Tk, q, kv);
  if (rc) return rc;
  if (!o || !ml) return fail(MMR_ERR_INVALID_ARG, "attention: null output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MMR_DTYPE_BF16 ? attention_core<bf16>(false, B, Tq, Tk, q, kv, kmask, o, ml, nullptr, nullptr, nullptr, nullptr, st)
                                 : attention_core<float>(false, B, Tq, Tk, q, kv, kmask, o, ml, nullptr, nullptr, nullptr, nullptr, st);
}

int mmr_attention_bwd(int dtype, int B, int Tq, int Tk, const void* q, const void* kv, const float* kmask, const void* o,
                      const float* ml, const void* d_o, void* dq, void* dkv, float* dvec, void* stream) {
  int rc = check_attention(dtype, B, Tq, Tk, q, kv);
  if (rc) return rc;
  if (!o || !ml || !d_o || !dq || !dkv || !dvec) return fail(MMR_ERR_INVALID_ARG, "attention: null pointer argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dtype == MMR_DTYPE_BF16
             ? attention_core<bf16>(true, B, Tq, Tk, q, kv, kmask, const_cast<void*>(o), const_cast<float*>(ml), d_o, dq, dkv, dvec, st)
             : attention_core<float>(true, B, Tq, Tk, q, kv, kmask, const_cast<void*>(o), const_cast<float*>(ml), d_o, dq, dkv, dvec, st);
}

int mmr_routing_stats_accumulate(const void* rc_raw, int rc_dtype, const float* rc_report, const float* prim_acts, int B,
                                 int K, float* sums, unsigned long long* count, void* stream) {
  if (!rc_raw || !prim_acts || !sums) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (B <= 0 || K < 1 || K > MMR_MAX_LABELS) return fail(MMR_ERR_INVALID_ARG, "bad batch / label count");
  if (rc_dtype != MMR_DTYPE_F32 && rc_dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_UNSUPPORTED, "rc_raw must be fp32 or bf16");
  RouteStatsArgs a; memset(&a, 0, sizeof(a));
  a.rc_raw = rc_raw; a.rc_bf16 = rc_dtype == MMR_DTYPE_BF16; a.rc_report = rc_report; a.prim_acts = prim_acts;
  a.B = B; a.K = K; a.sums = sums; a.count = count;
  route_stats_kernel<<<10 * K + 10, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  LAUNCH_OK("route_stats");
  return MMR_OK;
}

// ------------------------------------------------------------ route-input producer projections ---
struct ProjPlan {
  long long rows; int rows_pad, din, dout; bool ln, bias, bf16c, tc; int xdt;
  size_t ct, s_xh, s_stat, saved_bytes, f_w, scratch_fwd, b_dy, b_dxh, b_w, scratch_bwd;
};
static int proj_plan(const mmr_proj_dims* d, ProjPlan* P) {
  if (!d) return fail(MMR_ERR_INVALID_ARG, "null dims");
  if (d->rows <= 0 || d->rows > (1ll << 30)) return fail(MMR_ERR_INVALID_ARG, "producer projection: bad row count");
  if (d->d_in % 128 || d->d_in < 128 || d->d_in > 1024) return fail(MMR_ERR_UNSUPPORTED, "producer projection: d_in must be a multiple of 128 in [128, 1024]");
  if (d->d_out % 256 || d->d_out < 256 || d->d_out > 1024) return fail(MMR_ERR_UNSUPPORTED, "producer projection: d_out must be a multiple of 256 in [256, 1024]");
  if (d->x_dtype != MMR_DTYPE_F32 && d->x_dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_UNSUPPORTED, "producer projection: x must be fp32 or bf16");
  if (d->dtype != MMR_DTYPE_F32 && d->dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_INVALID_ARG, "producer projection: bad compute dtype");
  P->rows = d->rows; P->rows_pad = pad_seg((int)d->rows); P->din = d->d_in; P->dout = d->d_out;
  P->ln = d->has_ln != 0; P->bias = d->has_bias != 0; P->bf16c = d->dtype == MMR_DTYPE_BF16; P->xdt = d->x_dtype;
  P->tc = P->bf16c && d->gemm_engine != MMR_GEMM_SIMT;
  if (d->gemm_engine == MMR_GEMM_TC && !P->bf16c) return fail(MMR_ERR_UNSUPPORTED, "tcgen05 engine requires bf16");
  P->ct = P->bf16c ? 2 : 4;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align256(n); return r; };
  P->s_xh = take((size_t)P->rows_pad * P->din * P->ct);       // GEMM operand: LayerNorm(x) (or x) in the compute type
  P->s_stat = take((size_t)P->rows * 2 * 4);
  P->saved_bytes = o;
  o = 0;
  P->f_w = take((size_t)P->dout * P->din * P->ct);            // W in the compute type
  P->scratch_fwd = o;
  o = 0;
  P->b_dy = take((size_t)P->rows_pad * P->dout * P->ct);
  P->b_dxh = take((size_t)P->rows_pad * P->din * 4);
  P->b_w = take((size_t)P->dout * P->din * P->ct);            // W^T in the compute type ([d_in, d_out])
  P->scratch_bwd = o;
  return MMR_OK;
}

int mmr_producer_proj_sizes(const mmr_proj_dims* dims, size_t* saved_bytes, size_t* scratch_fwd_bytes, size_t* scratch_bwd_bytes) {
  ProjPlan P;
  int rc = proj_plan(dims, &P);
  if (rc) return rc;
  if (saved_bytes) *saved_bytes = P.saved_bytes;
  if (scratch_fwd_bytes) *scratch_fwd_bytes = P.scratch_fwd;
  if (scratch_bwd_bytes) *scratch_bwd_bytes = P.scratch_bwd;
  return MMR_OK;
}

extern "C++" {
template <class CT>
static int proj_fwd_t(const ProjPlan& P, const void* x, const float* ln_w, const float* ln_b, const float* W, const float* bias,
                      float* y, uint8_t* saved, uint8_t* scratch, cudaStream_t st) {
  CT* xh = reinterpret_cast<CT*>(saved + P.s_xh);
  ProjRowArgs a; memset(&a, 0, sizeof(a));
  a.x = x; a.rows = P.rows; a.rows_pad = P.rows_pad; a.D = P.din; a.has_ln = P.ln; a.gamma = ln_w; a.beta = ln_b;
  a.out = xh; a.stat = reinterpret_cast<float*>(saved + P.s_stat);
  const int blocks = (P.rows_pad + 7) / 8;
  if (P.xdt == MMR_DTYPE_F32) proj_rows_fwd_kernel<float, CT><<<blocks, 256, 0, st>>>(a);
  else proj_rows_fwd_kernel<bf16, CT><<<blocks, 256, 0, st>>>(a);
  LAUNCH_OK("proj_rows_fwd");
  GemmProblem g; memset(&g, 0, sizeof(g));
  g.segs = single_seg((int)P.rows, 1); g.segs.row0[1] = P.rows_pad;
  g.N = P.dout; g.K = P.din; g.A = xh; g.lda = P.din;
  if (P.tc) {
    CT* wb = reinterpret_cast<CT*>(scratch + P.f_w);
    PackJobs pj; memset(&pj, 0, sizeof(pj));
    pj.n = 1;
    pj.j[0] = PackJob{W, P.dout, P.din, P.din, nullptr, 1.0f, wb, P.din, nullptr, 0};
    pack_kernel<CT><<<dim3((P.dout / 32) * (P.din / 32), 1), 256, 0, st>>>(pj);
    LAUNCH_OK("proj pack");
    g.B = wb; g.ldb = P.din;
    tc::TcEpi e; memset(&e, 0, sizeof(e));
    e.bias = bias; e.out = y; e.ldo = P.dout; e.out_rows = (int)P.rows;
    ProfScope ps(PC_GEMM_TC, st);
    cudaError_t err = tc::launch_gemm_tc<tc::TEPI_BIAS_F32>(g, e, P.rows_pad, P.dout, st);
    if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 producer gemm: ") + cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    g.segs.row0[1] = (int)P.rows;       // the SIMT engine guards rows itself: no stores beyond y
    g.B = W; g.ldb = P.din;
    EpiParams e; memset(&e, 0, sizeof(e));
    e.bias = bias; e.out = y; e.ldo = P.dout;
    launch_gemm_simt<CT, float, EPI_BIAS_F32, CT>(g, e, st);
    LAUNCH_OK("producer gemm");
  }
  return MMR_OK;
}

}  // extern "C++"

int mmr_producer_proj_fwd(const mmr_proj_dims* dims, const void* x, const float* ln_w, const float* ln_b, const float* W,
                          const float* bias, float* y, void* saved, void* scratch, void* stream) {
  ProjPlan P;
  int rc = proj_plan(dims, &P);
  if (rc) return rc;
  if (!x || !W || !y || !saved || !scratch) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (P.ln && (!ln_w || !ln_b)) return fail(MMR_ERR_INVALID_ARG, "has_ln needs ln_w and ln_b");
  if (P.bias && !bias) return fail(MMR_ERR_INVALID_ARG, "has_bias needs bias");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.bf16c) return proj_fwd_t<bf16>(P, x, ln_w, ln_b, W, P.bias ? bias : nullptr, y, (uint8_t*)saved, (uint8_t*)scratch, st);
  return proj_fwd_t<float>(P, x, ln_w, ln_b, W, P.bias ? bias : nullptr, y, (uint8_t*)saved, (uint8_t*)scratch, st);
}

extern "C++" {
template <class CT>
static int proj_bwd_t(const ProjPlan& P, const void* x, const float* ln_w, const float* W, const float* dy, const uint8_t* saved,
                      uint8_t* scratch, float* dx, float* d_ln_w, float* d_ln_b, float* dW, float* dbias, cudaStream_t st) {
  const CT* xh = reinterpret_cast<const CT*>(saved + P.s_xh);
  CT* dyc = reinterpret_cast<CT*>(scratch + P.b_dy);
  float* dxh = reinterpret_cast<float*>(scratch + P.b_dxh);
  const bool need_dx = dx != nullptr || (P.ln && (d_ln_w || d_ln_b));
  {
    int blocks = (int)(((size_t)P.rows_pad * P.dout / 4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cast_rows_kernel<CT><<<blocks, 256, 0, st>>>(dy, P.rows, P.rows_pad, P.dout, dyc);
    LAUNCH_OK("cast dy");
  }
  float* dx_target = P.ln ? dxh : dx;       // without LayerNorm the data gradient is the GEMM output itself
  if (need_dx && dx_target) {               // dxh[rows, d_in] = dy[rows, d_out] * W[d_out, d_in]
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = single_seg((int)P.rows, 1); g.segs.row0[1] = P.rows_pad;
    g.N = P.din; g.K = P.dout; g.A = dyc; g.lda = P.dout;
    if (P.tc) {
      CT* wt = reinterpret_cast<CT*>(scratch + P.b_w);
      PackJobs pj; memset(&pj, 0, sizeof(pj));
      pj.n = 1;
      pj.j[0] = PackJob{W, P.dout, P.din, P.din, nullptr, 1.0f, nullptr, 0, wt, P.dout};
      pack_kernel<CT><<<dim3((P.dout / 32) * (P.din / 32), 1), 256, 0, st>>>(pj);
      LAUNCH_OK("proj pack T");
      g.B = wt; g.ldb = P.dout;
      tc::TcEpi e; memset(&e, 0, sizeof(e));
      e.out = dx_target; e.ldo = P.din; e.out_rows = P.ln ? P.rows_pad : (int)P.rows;
      ProfScope ps(PC_GEMM_TC, st);
      cudaError_t err = tc::launch_gemm_tc<tc::TEPI_F32>(g, e, P.rows_pad, P.din, st);
      if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 producer dgrad: ") + cudaGetErrorString(err));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      g.segs.row0[1] = (int)P.rows;
      g.B = W; g.ldb = P.din;
      EpiParams e; memset(&e, 0, sizeof(e));
      e.out = dx_target; e.ldo = P.din;
      launch_gemm_simt<CT, float, EPI_STORE_F32, CT, true>(g, e, st);
      LAUNCH_OK("producer dgrad");
    }
  }
  if (dW || (P.bias && dbias)) {            // dW[d_out, d_in] = dy^T xh, dbias = column sums of dy
    WgradProblem w; memset(&w, 0, sizeof(w));
    w.segs = single_seg((int)P.rows, 1); w.segs.row0[1] = P.rows_pad;
    w.dY = dyc; w.ldy = P.dout; w.X = xh; w.ldx = P.din; w.M = P.dout; w.N = P.din; w.out[0] = dW; w.ldo = P.din;
    bool fused_bias = false;
    if (P.tc && dW) {
      if (P.bias && dbias) { w.dbias[0] = dbias; w.colsum = 1; fused_bias = true; }
      ProfScope ps(PC_WGRAD_TC, st);
      cudaError_t err = tc::launch_wgrad_tc(w, P.rows_pad, P.rows_pad, st);
      if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 producer wgrad: ") + cudaGetErrorString(err));
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (dW) {
      launch_wgrad_simt<CT, CT>(w, st);
      LAUNCH_OK("producer wgrad");
    }
    if (P.bias && dbias && !fused_bias) {
      float* o1[6] = {dbias, nullptr, nullptr, nullptr, nullptr, nullptr};
      int rc = run_colsum<float>(single_seg((int)P.rows, 1), dy, P.dout, 0, P.dout, o1, 1.0f, st, "producer dbias");
      if (rc) return rc;
    }
  }
  if (P.ln && need_dx) {
    ProjRowArgs a; memset(&a, 0, sizeof(a));
    a.x = x; a.rows = P.rows; a.rows_pad = P.rows_pad; a.D = P.din; a.has_ln = 1; a.gamma = ln_w;
    a.stat = const_cast<float*>(reinterpret_cast<const float*>(saved + P.s_stat));
    a.dh = dxh; a.dx = dx; a.dgamma = d_ln_w; a.dbeta = d_ln_b;
    int blocks = (int)((P.rows + 63) / 64);
    if (blocks > 148 * 4) blocks = 148 * 4;
    const int rpb = (int)((P.rows + blocks - 1) / blocks);
    if (P.xdt == MMR_DTYPE_F32) proj_rows_bwd_kernel<float><<<blocks, 256, 0, st>>>(a, rpb);
    else proj_rows_bwd_kernel<bf16><<<blocks, 256, 0, st>>>(a, rpb);
    LAUNCH_OK("proj_rows_bwd");
  }
  return MMR_OK;
}

}  // extern "C++"

int mmr_producer_proj_bwd(const mmr_proj_dims* dims, const void* x, const float* ln_w, const float* W, const float* dy,
                          const void* saved, void* scratch, float* dx, float* d_ln_w, float* d_ln_b, float* dW, float* dbias,
                          void* stream) {
  ProjPlan P;
  int rc = proj_plan(dims, &P);
  if (rc) return rc;
  if (!x || !W || !dy || !saved || !scratch) return fail(MMR_ERR_INVALID_ARG, "null pointer argument");
  if (P.ln && !ln_w) return fail(MMR_ERR_INVALID_ARG, "has_ln needs ln_w");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.bf16c)
    return proj_bwd_t<bf16>(P, x, ln_w, W, dy, (const uint8_t*)saved, (uint8_t*)scratch, dx, d_ln_w, d_ln_b, dW, dbias, st);
  return proj_bwd_t<float>(P, x, ln_w, W, dy, (const uint8_t*)saved, (uint8_t*)scratch, dx, d_ln_w, d_ln_b, dW, dbias, st);
}

// ----------------------------------------------------------------------------- debug GEMM ---
int mmr_debug_gemm(int engine, int dtype, int trans, int M, int N, int K, const void* A, const void* B,
                   const float* bias, float* C, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N <= 0 || K <= 0) return fail(MMR_ERR_INVALID_ARG, "bad GEMM size");
  const bool use_tc = engine == MMR_GEMM_TC;
  if (use_tc && dtype != MMR_DTYPE_BF16) return fail(MMR_ERR_UNSUPPORTED, "tcgen05 engine requires bf16");
  if (!trans) {
    GemmProblem g; memset(&g, 0, sizeof(g));
    g.segs = single_seg(M, 1);
    g.N = N; g.K = K; g.A = A; g.lda = K; g.B = B; g.ldb = K;
    EpiParams e; memset(&e, 0, sizeof(e));
    e.bias = bias; e.out = C; e.ldo = N;
    if (use_tc) {
      if (K % 64 || N % 256) return fail(MMR_ERR_UNSUPPORTED, "tcgen05 debug gemm needs K%64==0, N%256==0");
      tc::TcEpi te; memset(&te, 0, sizeof(te));
      te.bias = bias; te.out = C; te.ldo = N;
      cudaError_t err = tc::launch_gemm_tc<tc::TEPI_BIAS_F32>(g, te, M, N, st);
      if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 gemm: ") + cudaGetErrorString(err));
    } else {
      if (K % 16 || N % 4) return fail(MMR_ERR_UNSUPPORTED, "simt debug gemm needs K%16==0, N%4==0");
      if (dtype == MMR_DTYPE_BF16) launch_gemm_simt<bf16, bf16, EPI_BIAS_F32, bf16>(g, e, st);
      else launch_gemm_simt<float, float, EPI_BIAS_F32, float>(g, e, st);
      LAUNCH_OK("debug gemm");
    }
  } else {   // C[M,N] += A[K,M]^T B[K,N]
    WgradProblem w; memset(&w, 0, sizeof(w));
    w.segs = single_seg(K, 1);
    w.segs.row0[1] = pad128(K);
    w.dY = A; w.ldy = M; w.X = B; w.ldx = N; w.M = M; w.N = N; w.out[0] = C; w.ldo = N;
    CUDA_OK(cudaMemsetAsync(C, 0, (size_t)M * N * 4, st));
    if (use_tc) {
      if (M % 128 || N % 256) return fail(MMR_ERR_UNSUPPORTED, "tcgen05 debug wgrad needs M%128==0, N%256==0");
      cudaError_t err = tc::launch_wgrad_tc(w, K, K, st);
      if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("tcgen05 wgrad: ") + cudaGetErrorString(err));
    } else {
      if (dtype == MMR_DTYPE_BF16) launch_wgrad_simt<bf16, bf16>(w, st);
      else launch_wgrad_simt<float, float>(w, st);
      LAUNCH_OK("debug wgrad");
    }
  }
  return MMR_OK;
}


/* Tuning hook: times `iters` back-to-back launches of the persistent tcgen05 GEMM  C[M,N] = A[M,K] B[N,K]^T with the
 * epilogue `op` (0 bias->bf16, 1 bias+relu+sign bits, 2 sign-bit mask, 4 fp32 out) using CUDA events on `stream`;
 * returns the average milliseconds per launch. */
int mmr_bench_gemm(int op, int M, int N, int K, const void* A, const void* B, const float* bias, void* C,
                   uint32_t* bits, int iters, float* ms_out, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M <= 0 || N % 256 || K % 64 || iters < 1) return fail(MMR_ERR_INVALID_ARG, "bad bench gemm size");
  GemmProblem g; memset(&g, 0, sizeof(g));
  g.segs = single_seg(M, 1);
  g.segs.row0[1] = pad128(M);
  g.N = N; g.K = K; g.A = A; g.lda = K; g.B = B; g.ldb = K;
  tc::TcEpi te; memset(&te, 0, sizeof(te));
  te.bias = bias; te.out = C; te.ldo = N; te.bits_in = bits; te.bits_out = bits; te.ld_bits = N / 32;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t err = cudaSuccess;
  for (int i = -1; i < iters && err == cudaSuccess; ++i) {
    if (i == 0) cudaEventRecord(e0, st);
    if (op == 0) err = tc::launch_gemm_tc<tc::TEPI_BIAS>(g, te, M, N, st);
    else if (op == 1) err = tc::launch_gemm_tc<tc::TEPI_BIAS_RELU_BITS>(g, te, M, N, st);
    else if (op == 2) err = tc::launch_gemm_tc<tc::TEPI_BITS_IN>(g, te, M, N, st);
    else err = tc::launch_gemm_tc<tc::TEPI_F32>(g, te, M, N, st);
  }
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (err != cudaSuccess) return fail(MMR_ERR_CUDA, std::string("bench gemm: ") + cudaGetErrorString(err));
  if (ms_out) *ms_out = ms / iters;
  return MMR_OK;
}


/* Tuning hook for the chained FFN kernel: mid[M,1024] = relu(A[M,256] B1[1024,256]^T + b1) (+ sign bits),
 * out[M,256] = mid B2[256,1024]^T + b2 (fwd = 1) or the backward pair (fwd = 0: bit mask, row mask). */
}  // extern "C"
