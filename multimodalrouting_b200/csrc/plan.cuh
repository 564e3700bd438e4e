// Host-side plan: row spaces, parameter indexing and buffer carving for the route-fusion path.
#pragma once
#include <string.h>

#include "mmr_common.cuh"

namespace mmr {

inline int pad128(int x) { return (x + 127) / 128 * 128; }
// row-space segments are 256-row aligned: CTA pairs (tcgen05 cta_group::2) own 256-row tiles
inline int pad_seg(int x) { return (x + 255) / 256 * 256; }
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// Index of parameter tensors in the host pointer table (MULTModel state_dict order).
struct ParamIndex {
  int layers;
  int proj(int mod) const { return mod; }                                  // proj_{l,n,i}.weight
  int uni_ln(int mod, int wb) const { return 3 + mod * 2 + wb; }           // trans_{l,n,i}.layer_norm
  int enc_base(int dir) const { return 9 + dir * (layers * 12 + 2); }
  // per-layer slots: 0 in_proj_weight 1 in_proj_bias 2 out_proj.weight 3 out_proj.bias 4 fc1.weight
  // 5 fc1.bias 6 fc2.weight 7 fc2.bias 8 ln0.weight 9 ln0.bias 10 ln1.weight 11 ln1.bias
  int layer(int dir, int l, int slot) const { return enc_base(dir) + l * 12 + slot; }
  int enc_ln(int dir, int wb) const { return enc_base(dir) + layers * 12 + wb; }
  int pair(int p, int wb) const { return 9 + NDIR * (layers * 12 + 2) + p * 2 + wb; }   // ln, li, ni
  int final_lni(int wb) const { return pair(3, wb); }
  int count() const { return pair(4, 0); }
};

struct Plan {
  mmr_fusion_dims dims;
  int B, L;
  int T[NMOD];            // tokens per modality
  int din[NMOD];          // input feature dims
  Segs mod;               // modality row space  (rows = B*T[m])
  Segs q;                 // query row space, one segment per direction (rows = B*T[qmod])
  Segs kv;                // key/value row space, one segment per direction (rows = B*T[kmod])
  int MM, MQ, MK;         // padded totals
  bool bf16;              // compute type
  bool tc;                // tcgen05 engine
  size_t ct;              // sizeof(compute type)

  // ---- packed weights (device) ------------------------------------------------------------
  // CT matrices stacked over directions so one grouped GEMM covers all six.
  size_t o_wq, o_wo, o_w1, o_w2;         // [L][6*N][K] forward weights (wq pre-scaled by hd^-0.5)
  size_t o_wqT, o_woT, o_w1T, o_w2T;     // [L][6*K][N] transposes for data gradients
  size_t o_wkv, o_wkvT;                  // [6*L*512][256] (LN0 gamma folded), [6*256][L*512]
  size_t o_bq, o_bo, o_b1, o_b2, o_bkv;  // fp32 biases: [L][6*N]; bkv [6*L*512] (LN0 beta folded)
  size_t packed_bytes;

  // ---- saved for backward -----------------------------------------------------------------
  size_t s_xh, s_rstd_e;                  // CT [MM,256] normalised embeddings; fp32 [MM]
  size_t s_maskq;                         // fp32 [MQ] query keep-mask per q row
  size_t s_xin, s_stat0, s_h0, s_qb, s_o, s_ml, s_x1, s_stat1, s_h1, s_f;   // per layer (stride *_l)
  size_t l_xin, l_stat, l_ct256, l_ml, l_f;
  size_t s_statf;                         // fp32 [MQ,2]
  size_t s_bits, l_bits;                  // u32 [L][MQ,32] ReLU sign bits of fc1 (tcgen05 engine)
  size_t s_kv;                            // CT [MK, L*512]
  size_t s_epair;                         // fp32 [3,B,256] pair embeddings eLN,eLI,eNI
  size_t s_routes;                        // fp32 [10,B,256] copy of the outputs (pair/trimodal bwd)
  size_t s_cnt;                           // fp32 [3,B] valid-token counts (clamped >= 1)
  // row plan of the packed query space (int32; written by the rowplan_* kernels in the forward, read by every q-space kernel of
  // both passes): nv[8] | per query modality m: poff[B+1], tokrow[B*T_m] (token -> row inside the segment, -1 = padded),
  // rowpat[B*T_m] (row -> patient)
  size_t s_plan;
  size_t p_poff[NMOD], p_tokrow[NMOD], p_rowpat[NMOD];   // int32 offsets from s_plan
  size_t saved_bytes;

  // ---- forward scratch ----------------------------------------------------------------------
  size_t f_p;                             // fp32 [MM,256] projected inputs (when din != 256)
  size_t f_y;                             // fp32 [MQ,256] final-LN outputs before pooling
  size_t f_u;                             // fp32 [MM,256] unimodal encoder outputs
  size_t f_delta;                         // CT [MQ,256] out_proj / fc2 outputs before the residual add
  size_t scratch_fwd_bytes;

  // ---- backward scratch ---------------------------------------------------------------------
  size_t b_g, b_g1;                       // fp32 [MQ,256] residual-stream gradients (ping/pong)
  size_t b_gc, b_gc2;                     // CT [MQ,256] copy of the current gradient (ping/pong: the weight-gradient
                                          // stream may still read one while the chain writes the other)
  size_t b_df;                            // CT [MQ,1024]
  size_t b_dh;                            // CT [MQ,256]   (dH1 / dH0)
  size_t b_do, b_dq;                      // CT [MQ,256]
  size_t b_dkv;                           // CT [MK, L*512]
  size_t b_dxh;                           // fp32 [MK,256]
  size_t b_dp;                            // fp32 [MM,256] gradient wrt projected inputs
  size_t b_dwq, b_dwkv;                   // fp32 [L][6][256*256], [6][L*512*256]   (packed-weight grads)
  size_t b_dbq, b_dbkv;                   // fp32 [L][6*256], [6*L*512]
  size_t b_depair;                        // fp32 [3,B,256]
  size_t b_dzcat;                         // fp32 [3,B,512] gradient wrt the concatenated pair operands
  size_t b_dvec;                          // fp32 [MQ,8] attention row statistic sum_c dO*O
  size_t b_zero_begin, b_zero_end;        // region that must be zeroed before accumulation
  size_t scratch_bwd_bytes;
};

inline bool build_plan(const mmr_fusion_dims* d, Plan* p, const char** why) {
  memset(p, 0, sizeof(*p));
  if (!d) { *why = "dims is NULL"; return false; }
  if (d->B <= 0 || d->TL <= 0 || d->TN <= 0 || d->TI <= 0) { *why = "B and token counts must be positive"; return false; }
  if (d->layers < 1 || d->layers > MMR_MAX_LAYERS) { *why = "layers must be in [1,8]"; return false; }
  if (d->dL % 16 || d->dN % 16 || d->dI % 16 || d->dL <= 0 || d->dN <= 0 || d->dI <= 0) { *why = "input dims must be positive multiples of 16"; return false; }
  if (d->dtype != MMR_DTYPE_F32 && d->dtype != MMR_DTYPE_BF16) { *why = "unknown dtype"; return false; }
  if ((long long)d->B * (d->TL + d->TN + d->TI) > (1ll << 27)) { *why = "batch too large for 32-bit row indexing"; return false; }
  p->dims = *d;
  p->B = d->B; p->L = d->layers;
  p->T[0] = d->TL; p->T[1] = d->TN; p->T[2] = d->TI;
  p->din[0] = d->dL; p->din[1] = d->dN; p->din[2] = d->dI;
  p->bf16 = d->dtype == MMR_DTYPE_BF16;
  p->tc = p->bf16 && d->gemm_engine != MMR_GEMM_SIMT;
  if (!p->bf16 && d->gemm_engine == MMR_GEMM_TC) { *why = "tcgen05 engine requires bf16"; return false; }
  p->ct = p->bf16 ? 2 : 4;
  p->mod.n = NMOD;
  int r = 0;
  for (int m = 0; m < NMOD; ++m) { p->mod.row0[m] = r; p->mod.rows[m] = p->B * p->T[m]; p->mod.T[m] = p->T[m]; r += pad_seg(p->mod.rows[m]); }
  p->mod.row0[NMOD] = r; p->MM = r;
  p->q.n = NDIR; p->kv.n = NDIR;
  int rq = 0, rk = 0;
  for (int dd = 0; dd < NDIR; ++dd) {
    const int qm = dir_qmod(dd), km = dir_kmod(dd);
    p->q.row0[dd] = rq; p->q.rows[dd] = p->B * p->T[qm]; p->q.T[dd] = p->T[qm]; rq += pad_seg(p->q.rows[dd]);
    p->kv.row0[dd] = rk; p->kv.rows[dd] = p->B * p->T[km]; p->kv.T[dd] = p->T[km]; rk += pad_seg(p->kv.rows[dd]);
  }
  p->q.row0[NDIR] = rq; p->MQ = rq;
  p->kv.row0[NDIR] = rk; p->MK = rk;

  const size_t ct = p->ct, L = p->L;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align256(o + bytes); return at; };
  p->o_wq = take(L * 6 * D * D * ct);   p->o_wo = take(L * 6 * D * D * ct);
  p->o_w1 = take(L * 6 * FF * D * ct);  p->o_w2 = take(L * 6 * D * FF * ct);
  p->o_wqT = take(L * 6 * D * D * ct);  p->o_woT = take(L * 6 * D * D * ct);
  p->o_w1T = take(L * 6 * D * FF * ct); p->o_w2T = take(L * 6 * FF * D * ct);
  p->o_wkv = take(6 * L * 2 * D * D * ct); p->o_wkvT = take(6 * D * L * 2 * D * ct);
  p->o_bq = take(L * 6 * D * 4); p->o_bo = take(L * 6 * D * 4); p->o_b1 = take(L * 6 * FF * 4);
  p->o_b2 = take(L * 6 * D * 4); p->o_bkv = take(6 * L * 2 * D * 4);
  p->packed_bytes = o;

  o = 0;
  const size_t MM = p->MM, MQ = p->MQ, MK = p->MK, B = p->B;
  p->s_xh = take(MM * D * ct); p->s_rstd_e = take(MM * 4); p->s_maskq = take(MQ * 4);
  p->l_xin = align256(MQ * D * 4); p->l_stat = align256(MQ * 2 * 4); p->l_ct256 = align256(MQ * D * ct);
  p->l_ml = align256(MQ * H * 2 * 4); p->l_f = align256(MQ * FF * ct);
  p->s_xin = take((L + 1) * p->l_xin);
  p->s_stat0 = take(L * p->l_stat); p->s_h0 = take(L * p->l_ct256); p->s_qb = take(L * p->l_ct256);
  p->s_o = take(L * p->l_ct256); p->s_ml = take(L * p->l_ml); p->s_x1 = take(L * p->l_xin);
  p->s_stat1 = take(L * p->l_stat); p->s_h1 = take(L * p->l_ct256); p->s_f = take(L * p->l_f);
  p->s_statf = take(MQ * 2 * 4);
  p->l_bits = align256(MQ * (FF / 32) * 4); p->s_bits = take(L * p->l_bits);
  p->s_kv = take(MK * L * 2 * D * ct);
  p->s_epair = take(3 * B * D * 4); p->s_routes = take(10 * B * D * 4); p->s_cnt = take(3 * B * 4);
  {
    size_t ints = 8;
    for (int m = 0; m < NMOD; ++m) {
      p->p_poff[m] = ints; ints += (size_t)B + 1;
      p->p_tokrow[m] = ints; ints += (size_t)B * p->T[m];
      p->p_rowpat[m] = ints; ints += (size_t)B * p->T[m];
    }
    p->s_plan = take(ints * 4);
  }
  p->saved_bytes = o;

  o = 0;
  p->f_p = take(MM * D * 4); p->f_y = take(MQ * D * 4); p->f_u = take(MM * D * 4); p->f_delta = take(MQ * D * ct);
  p->scratch_fwd_bytes = o;

  o = 0;
  p->b_g = take(MQ * D * 4); p->b_g1 = take(MQ * D * 4); p->b_gc = take(MQ * D * ct); p->b_gc2 = take(MQ * D * ct);
  p->b_df = take(MQ * FF * ct); p->b_dh = take(MQ * D * ct); p->b_do = take(MQ * D * ct); p->b_dq = take(MQ * D * ct);
  p->b_dkv = take(MK * L * 2 * D * ct);
  p->b_dp = take(MM * D * 4); p->b_dvec = take(MQ * H * 4);
  p->b_zero_begin = o;
  p->b_dxh = take(MK * D * 4);     // zeroed: only because pad rows are read by the modality reduction
  p->b_dwq = take(L * 6 * D * D * 4); p->b_dwkv = take(6 * L * 2 * D * D * 4);
  p->b_dbq = take(L * 6 * D * 4); p->b_dbkv = take(6 * L * 2 * D * 4);
  p->b_depair = take(3 * B * D * 4); p->b_dzcat = take(3 * B * 512 * 4);
  p->b_zero_end = o;
  p->scratch_bwd_bytes = o;
  return true;
}

}  // namespace mmr
