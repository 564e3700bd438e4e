/* mmr_b200.h -- C ABI of the B200-native route-fusion + capsule-routing hot path.
 *
 * This is the drop-in boundary for ONE path of AI-for-Health-Data/MultimodalRouting:
 *   MULTModel.forward                     MIMIC-IV/MortModel/Paired_Cross_Attention/mult_model.py:116-193
 *   TransformerEncoder(.Layer).forward    .../transformer.py:56-115,149-216
 *   MultiheadAttention.forward            MIMIC-IV/PhenoModel/Paired_Cross_Attention/multihead_attention.py:48-148
 *   RoutePrimaryProjector.forward         .../routing_and_heads.py:111-121
 *   forward_capsule_from_route_dict       .../routing_and_heads.py:271-369
 *   CapsuleMortalityHead.forward          Mort .../routing_and_heads.py:194-268, Pheno :194-272
 *   CapsuleFC.forward                     .../capsule_layers.py:75-117
 * The reference has no FFI of its own (it is eager PyTorch); the Python nn.Modules in
 * multimodalrouting_b200/ bind these entry points through ctypes + torch.library.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless named host_*;
 *   - the library never allocates or frees device memory and keeps no global mutable state
 *     (apart from a thread-local last-error string); the caller owns inputs, outputs, packed
 *     weights, saved-for-backward and scratch buffers (sizes from mmr_fusion_sizes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises;
 *   - every function returns 0 on success, non-zero otherwise (mmr_last_error_string()).
 *   - fixed model geometry (reference hyper-parameters): d=256, heads=8,
 *     head_dim=32, ffn=1024, 10 routes, pc_dim=32, mc_caps_dim=64.
 */
#ifndef MMR_B200_H
#define MMR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMR_OK 0
#define MMR_ERR_INVALID_ARG 1
#define MMR_ERR_UNSUPPORTED 2
#define MMR_ERR_CUDA 3

#define MMR_DTYPE_F32 0   /* fp32 SIMT arithmetic everywhere (parity mode, 1e-4)          */
#define MMR_DTYPE_BF16 1  /* bf16 operands, fp32 accumulate / residual / softmax / LayerNorm */

#define MMR_GEMM_AUTO 0   /* tcgen05 for bf16, SIMT for fp32 */
#define MMR_GEMM_SIMT 1   /* debug: CUDA-core GEMM for every dtype */
#define MMR_GEMM_TC 2

#define MMR_VARIANT_MORT 0   /* routing uses act = route_mask, d = sum_r R*pose           */
#define MMR_VARIANT_PHENO 1  /* routing uses act = alpha,      d = sum_r R*alpha*pose     */

#define MMR_D 256
#define MMR_HEADS 8
#define MMR_HEAD_DIM 32
#define MMR_FFN 1024
#define MMR_ROUTES 10
#define MMR_PC_DIM 32
#define MMR_MC_DIM 64
#define MMR_MAX_LAYERS 8
#define MMR_MAX_LABELS 32

typedef struct mmr_fusion_dims {
  int32_t B;             /* patients */
  int32_t TL, TN, TI;    /* tokens per modality */
  int32_t dL, dN, dI;    /* input feature dims; != 256 enables the Conv1d(k=1) projection */
  int32_t layers;        /* cross-modal encoder depth (reference: 4) */
  int32_t dtype;         /* MMR_DTYPE_* */
  int32_t gemm_engine;   /* MMR_GEMM_* */
} mmr_fusion_dims;

/* Number of parameter tensors of MULTModel in state_dict order (mult_model.py:30-57):
 * proj_{l,n,i}.weight, trans_{l,n,i}.layer_norm.{weight,bias}, then per cross encoder
 * (l_with_n, l_with_i, n_with_l, n_with_i, i_with_l, i_with_n) per layer
 * {in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias, fc1.weight, fc1.bias,
 *  fc2.weight, fc2.bias, layer_norms.0.{weight,bias}, layer_norms.1.{weight,bias}} and the
 * encoder's final layer_norm.{weight,bias}; then proj_pair_{ln,li,ni}.{weight,bias},
 * final_lni.{weight,bias}.  All fp32, contiguous. */
int mmr_fusion_num_params(const mmr_fusion_dims* dims);

/* Buffer sizes in bytes.  `saved` carries forward->backward state. */
int mmr_fusion_sizes(const mmr_fusion_dims* dims, size_t* packed_bytes, size_t* saved_bytes,
                     size_t* scratch_fwd_bytes, size_t* scratch_bwd_bytes);

/* Replaces MULTModel.forward (mult_model.py:116-193).
 * host_params: host array of mmr_fusion_num_params() device pointers (order above).
 * x_*: fp32 [B,T,d_*]; m*: fp32 [B,T] (1 keep / 0 pad) or NULL; pos_table: fp32 [maxT,256]
 * int-truncated sinusoid rows for positions 1..maxT (position_embedding.py:68-117).
 * routes_out: fp32 [10,B,256] in ROUTES order L,N,I,LN,NL,LI,IL,NI,IN,LNI. */
int mmr_route_fusion_fwd(const mmr_fusion_dims* dims, const void* const* host_params,
                         const float* x_l, const float* x_n, const float* x_i,
                         const float* mL, const float* mN, const float* mI,
                         const float* pos_table, void* packed, void* saved, void* scratch,
                         float* routes_out, void* stream);

/* The same forward in two steps, for callers whose weights do not change between calls (evaluation, several forward
 * passes per optimizer step): mmr_fusion_pack_weights fills `packed` (mmr_fusion_sizes' packed_bytes; depends on layers and
 * dtype only) from the fp32 parameters -- once per parameter version -- and mmr_route_fusion_fwd_packed runs the forward
 * from it without re-packing.  `packed` is also what the backward entry points read. */
int mmr_fusion_pack_weights(const mmr_fusion_dims* dims, const void* const* host_params, void* packed, void* stream);
int mmr_route_fusion_fwd_packed(const mmr_fusion_dims* dims, const void* const* host_params,
                                const float* x_l, const float* x_n, const float* x_i,
                                const float* mL, const float* mN, const float* mI,
                                const float* pos_table, const void* packed, void* saved, void* scratch,
                                float* routes_out, void* stream);

/* Backward of the above.  d_routes: fp32 [10,B,256].  host_param_grads: host array of device
 * pointers (same order; NULL entries skipped) to ZERO-INITIALISED fp32 gradient tensors; the
 * gradients are accumulated into them.  dx_*: fp32 [B,T,d_*] outputs or NULL. */
int mmr_route_fusion_bwd(const mmr_fusion_dims* dims, const void* const* host_params,
                         const float* x_l, const float* x_n, const float* x_i,
                         const float* mL, const float* mN, const float* mI,
                         const void* packed, const void* saved, void* scratch,
                         const float* d_routes, void* const* host_param_grads,
                         float* dx_l, float* dx_n, float* dx_i, void* stream);

/* Same as mmr_route_fusion_bwd; additionally records host_layer_events[l] (cudaEvent_t, NULL entries skipped) on
 * `stream` once the out_proj / fc1 / fc2 weight+bias and layer_norms.1 gradients of layer l (all six encoders) are
 * final, so a data-parallel caller can all-reduce that part of the gradient buffer while the backward continues. */
int mmr_route_fusion_bwd_events(const mmr_fusion_dims* dims, const void* const* host_params,
                                const float* x_l, const float* x_n, const float* x_i,
                                const float* mL, const float* mN, const float* mI,
                                const void* packed, const void* saved, void* scratch,
                                const float* d_routes, void* const* host_param_grads,
                                float* dx_l, float* dx_n, float* dx_i, void* stream,
                                void* const* host_layer_events);

/* Most general form.  host_layer_events: as above (may be NULL).  side_stream (may be NULL): a second cudaStream_t on
 * which the weight-gradient kernels run next to the data-gradient chain of `stream`; host_sync_events then holds
 * MMR_BWD_SYNC_EVENTS caller-owned cudaEvent_t (created with cudaEventDisableTiming) that the library records / waits on
 * to order the two streams.  The side stream is forked from and joined back into `stream` inside the call, so the
 * call is CUDA-graph capturable and everything is complete in `stream` order when it returns.  Ignored (single stream)
 * when host_layer_events is given. */
#define MMR_BWD_SYNC_EVENTS 9
int mmr_route_fusion_bwd_ex(const mmr_fusion_dims* dims, const void* const* host_params,
                            const float* x_l, const float* x_n, const float* x_i,
                            const float* mL, const float* mN, const float* mI,
                            const void* packed, const void* saved, void* scratch,
                            const float* d_routes, void* const* host_param_grads,
                            float* dx_l, float* dx_n, float* dx_i, void* stream,
                            void* const* host_layer_events, void* side_stream, void* const* host_sync_events);

typedef struct mmr_routing_dims {
  int32_t B;               /* patients */
  int32_t K;               /* labels: 2 (mortality) / 25 (phenotypes); <= MMR_MAX_LABELS */
  int32_t variant;         /* MMR_VARIANT_* */
  int32_t num_routing;     /* agreement iterations (reference: 3) */
  int32_t detach_priors;   /* routing_and_heads.py:352 */
  int32_t from_poses;      /* 0: run projector on route embeddings; 1: poses/acts given (head only) */
  float act_temperature;   /* routing_and_heads.py:330-339 (only applied when route_mask != NULL) */
  float prior_floor, prior_ceiling;   /* 0.02 / 0.98 (env_config.py:157-158) */
  int64_t emb_route_stride;   /* elements between routes of one patient in route_embs */
  int64_t emb_batch_stride;   /* elements between patients */
  int32_t vote_dtype;         /* MMR_DTYPE_F32: votes held in fp32 (parity mode); MMR_DTYPE_BF16: the reduced-
                                 precision mode -- votes held as saturating fp16 (11-bit mantissa, tighter than the
                                 reference's bf16 autocast einsums); all accumulation is fp32 in both modes */
  int32_t reserved;
} mmr_routing_dims;

typedef struct mmr_routing_params {
  const float* proj_w[MMR_ROUTES];   /* [33,256] each, ROUTES order */
  const float* proj_b[MMR_ROUTES];   /* [33] */
  const float* caps_w;               /* [10,32,K,64] */
  const float* pose_to_mc;           /* [64,32] */
  const float* embedding;            /* [K,64] */
  const float* bias;                 /* [K] */
  /* optional (NULL: CUDA-core contraction): fp16 copies written by mmr_routing_pack_weights, read by the tensor-core
   * projector / vote contraction of the reduced-precision mode (vote_dtype = MMR_DTYPE_BF16) */
  const void* caps_wt_f16;          /* [10, K*64, 32]  (transposed: forward vote contraction) */
  const void* caps_w_f16;           /* [10, 32, K*64]  (original layout: backward of the vote contraction) */
  const void* proj_w_f16;           /* [10, 40, 256], rows >= 33 zero */
} mmr_routing_params;

typedef struct mmr_routing_grads {   /* zero-initialised fp32 accumulators; NULL entries skipped */
  float* proj_w[MMR_ROUTES];
  float* proj_b[MMR_ROUTES];
  float* caps_w;
  float* pose_to_mc;
  float* embedding;
  float* bias;
} mmr_routing_grads;

size_t mmr_routing_scratch_bytes(const mmr_routing_dims* dims);
/* scratch of mmr_capsule_routing_fwd_ex (projector outputs, head matrix and the fp16 votes of the batch: 20 * K * 64 bytes
 * per patient + ~1.4 KB) */
size_t mmr_routing_fwd_scratch_bytes(const mmr_routing_dims* dims);

/* Packs capsule.w (and, when params->proj_w[0] != NULL and proj_w_f16 != NULL, the projector weights) into the fp16
 * layouts above.  caps_wt_f16 and caps_w_f16 need 10*K*64*32*2 bytes each, proj_w_f16 10*40*256*2 bytes (16-byte aligned, caller-owned).
 * Call once per parameter version (the Python op does it per forward and hands the buffers to the backward). */
int mmr_routing_pack_weights(const mmr_routing_params* params, int K, void* caps_wt_f16, void* caps_w_f16,
                             void* proj_w_f16, void* stream);

/* Replaces forward_capsule_from_route_dict / CapsuleMortalityHead.forward.
 * route_embs: fp32, element (r,b,c) at r*emb_route_stride + b*emb_batch_stride + c.
 * When from_poses=1: poses_in fp32 [B,10,32], acts_in fp32 [B,10] are used instead (no clamp /
 * temperature, exactly CapsuleMortalityHead.forward); acts_override fp32 [B,10] or NULL.
 * route_mask: fp32 [B,10] or NULL.  Outputs fp32: logits [B,K], alpha [B,10], R [B,10,K],
 * poses_out [B,10,32] and acts_out [B,10] (projector outputs; may be NULL). */
int mmr_capsule_routing_fwd(const mmr_routing_dims* dims, const mmr_routing_params* params,
                            const float* route_embs, const float* poses_in, const float* acts_in,
                            const float* acts_override, const float* route_mask,
                            float* logits, float* alpha, float* R, float* poses_out,
                            float* acts_out, void* stream);

/* Same call with a caller-owned scratch buffer (mmr_routing_fwd_scratch_bytes, 256-byte aligned; NULL = the call above).
 * With scratch, vote_dtype = MMR_DTYPE_BF16, the fp16 weight copies and num_routing <= 3 the work is split into
 * projector / vote GEMM launches over 16-patient tiles and an agreement kernel that gives every patient its own 1-4 warps
 * (csrc/routing_split.cuh); the backward takes the same path on its own from mmr_routing_scratch_bytes.  MMR_RT_SPLIT=0
 * in the environment keeps both on the tile-of-4 kernels. */
int mmr_capsule_routing_fwd_ex(const mmr_routing_dims* dims, const mmr_routing_params* params,
                               const float* route_embs, const float* poses_in, const float* acts_in,
                               const float* acts_override, const float* route_mask,
                               float* logits, float* alpha, float* R, float* poses_out,
                               float* acts_out, void* scratch, void* stream);

/* Backward: d_logits [B,K], d_R [B,10,K] or NULL.  d_route_embs uses the same strides as
 * route_embs; d_poses [B,10,32] / d_acts [B,10] are written when from_poses=1.  With from_poses=0 and acts_override
 * given, d_acts (may be NULL) receives the gradient wrt acts_override [B,10] (routing_and_heads.py:314: the prior chain
 * ends at the override instead of the projector's activation logit). */
int mmr_capsule_routing_bwd(const mmr_routing_dims* dims, const mmr_routing_params* params,
                            const float* route_embs, const float* poses_in, const float* acts_in,
                            const float* acts_override, const float* route_mask,
                            const float* d_logits, const float* d_R, void* scratch,
                            const mmr_routing_grads* grads, float* d_route_embs, float* d_poses,
                            float* d_acts, void* stream);

/* Same call; fwd_scratch (may be NULL) is the scratch buffer of the mmr_capsule_routing_fwd_ex call on the SAME inputs and
 * parameters, untouched since: the split path then reuses the projector outputs and votes instead of recomputing them. */
int mmr_capsule_routing_bwd_ex(const mmr_routing_dims* dims, const mmr_routing_params* params,
                               const float* route_embs, const float* poses_in, const float* acts_in,
                               const float* acts_override, const float* route_mask,
                               const float* d_logits, const float* d_R, void* scratch,
                               const mmr_routing_grads* grads, float* d_route_embs, float* d_poses,
                               float* d_acts, const void* fwd_scratch, void* stream);

/* Per-patient multi-head attention core (8 heads x 32) on the path's attention kernels, for ONE pair of streams -- what the
 * attention-fusion variants of the Partial/ model (`CrossAttentionFusion`, `TriTokenAttentionFusion`:
 * MIMIC-IV/PhenoModel/Partial/Cross_Attention/routing_and_heads.py:103-206) need between their projections:
 *   o[b, i, h*32:(h+1)*32] = sum_j softmax_j(q_h[b,i] . k_h[b,j] + keymask[b,j]) v_h[b,j]
 * q: [B*Tq, 256], already scaled by head_dim^-1/2; kv: [B*Tk, 512] = K | V per row; kmask: fp32 [B, Tk] (>= 0.5 keep) or NULL;
 * o: [B*Tq, 256]; ml: fp32 [B*Tq, 8, 2] (row max, 1 / row sum), saved for the backward.  dtype = MMR_DTYPE_BF16 (bf16 operands,
 * mma.sync tiles, scores rounded to bf16 before an fp32 softmax like the reference's bmm under autocast) or MMR_DTYPE_F32.
 * Key masking follows the hot path: padded keys get finfo(bf16).min, so a patient whose keys are all padded
 * attends uniformly instead of producing NaN; the fusion modules zero such patients afterwards, as the reference intends. */
int mmr_attention_fwd(int dtype, int B, int Tq, int Tk, const void* q, const void* kv, const float* kmask, void* o, float* ml,
                      void* stream);
/* Backward: d_o [B*Tq,256] -> dq [B*Tq,256], dkv [B*Tk,512]; dvec: fp32 [B*Tq, 8] scratch. */
int mmr_attention_bwd(int dtype, int B, int Tq, int Tk, const void* q, const void* kv, const float* kmask, const void* o,
                      const float* ml, const void* d_o, void* dq, void* dkv, float* dvec, void* stream);

/* Replaces RoutePrimaryProjector.forward alone (routing_and_heads.py:111-121) for callers that use the projector outside
 * forward_capsule_from_route_dict: poses [B,10,32] = (W_r e_r + b_r)[:32], acts [B,10] = sigmoid((W_r e_r + b_r)[32]).
 * route_embs / strides as above; only params->proj_w / proj_b are read.  fp32. */
int mmr_projector_fwd(const mmr_routing_params* params, const float* route_embs, int64_t emb_route_stride,
                      int64_t emb_batch_stride, int B, float* poses, float* acts, void* stream);
/* Backward of the above: d_poses [B,10,32] / d_acts [B,10] (either may be NULL = zero); scratch: B*330*4 bytes;
 * accumulates into grads->proj_w / proj_b (zero-initialised, NULL entries skipped; other fields ignored) and writes
 * d_route_embs (same strides as route_embs; may be NULL). */
int mmr_projector_bwd(const mmr_routing_params* params, const float* route_embs, int64_t emb_route_stride,
                      int64_t emb_batch_stride, int B, const float* d_poses, const float* d_acts, void* scratch,
                      const mmr_routing_grads* grads, float* d_route_embs, void* stream);

/* ---- the steps either side of the hot path ---------------------------
 *
 * Producer epilogue: replaces _clamp_norm + _safe_tensor + .float() of the encoder outputs
 * (MortModel/Paired_Cross_Attention/main.py:1772-1796, mode 0) or the nan_to_num-only variant of
 * PhenoModel/Paired_Cross_Attention/main.py:1445-1460 (mode 1).  x: [rows, D] of in_dtype (MMR_DTYPE_F32 /
 * MMR_DTYPE_BF16 / MMR_DTYPE_F16), y: fp32 [rows, D]; D % 4 == 0, D <= 1024.
 *   mode 0: y = nan_to_num(x * min(1, max_norm / (||x||_2 + 1e-6)), nan=0, posinf=1e4, neginf=-1e4)
 *   mode 1: y = nan_to_num(x, 0, 0, 0)
 * nonfinite (device, may be NULL) is incremented by the number of entries nan_to_num replaced. */
#define MMR_DTYPE_F16 2
int mmr_sanitize_rows_fwd(const void* x, int in_dtype, float* y, int64_t rows, int D, int mode, float max_norm,
                          unsigned long long* nonfinite, void* stream);
/* dx = d(y)/d(x)^T dy for rows of finite inputs; rows holding NaN/Inf get dx = 0 (the reference produces NaN there
 * and skips the optimizer step). */
int mmr_sanitize_rows_bwd(const void* x, int in_dtype, const float* dy, float* dx, int64_t rows, int D, int mode,
                          float max_norm, void* stream);

/* Route-input projections between the modality encoders and the hot path:
 *   BioClinicalBERT chunk projection  Sequential(LayerNorm(768), Linear(768 -> 256, bias=False))
 *                                     MIMIC-IV/MortModel/Paired_Cross_Attention/encoders.py:289-293, 472-475   (has_ln = 1)
 *   CXR token projection              Linear(512 -> 256, bias=False)              encoders.py:620, 747-749     (has_ln = 0)
 * y[rows, d_out] (fp32) = LN?(x[rows, d_in]) W^T (+ bias).  dtype = MMR_DTYPE_BF16: LayerNorm in fp32, GEMM operands bf16 on
 * the tcgen05 engine (the reference's autocast flow); MMR_DTYPE_F32: fp32 SIMT GEMM (parity mode).  x is fp32 or bf16
 * (x_dtype), so an encoder that already emits bf16 is consumed without a cast pass.  W: fp32 [d_out, d_in]; d_in % 128 == 0,
 * d_in <= 1024; d_out % 256 == 0.  saved (forward -> backward) and scratch sizes from mmr_producer_proj_sizes. */
typedef struct mmr_proj_dims {
  int64_t rows;
  int32_t d_in, d_out;
  int32_t has_ln, has_bias;
  int32_t x_dtype;          /* MMR_DTYPE_F32 / MMR_DTYPE_BF16 of x */
  int32_t dtype;            /* compute dtype, MMR_DTYPE_* */
  int32_t gemm_engine;      /* MMR_GEMM_* */
  int32_t reserved;
} mmr_proj_dims;
int mmr_producer_proj_sizes(const mmr_proj_dims* dims, size_t* saved_bytes, size_t* scratch_fwd_bytes,
                            size_t* scratch_bwd_bytes);
int mmr_producer_proj_fwd(const mmr_proj_dims* dims, const void* x, const float* ln_w, const float* ln_b, const float* W,
                          const float* bias, float* y, void* saved, void* scratch, void* stream);
/* dy: fp32 [rows, d_out].  Outputs (each may be NULL): dx fp32 [rows, d_in]; d_ln_w / d_ln_b fp32 [d_in], dW fp32
 * [d_out, d_in], dbias fp32 [d_out] -- ZERO-INITIALISED accumulators. */
int mmr_producer_proj_bwd(const mmr_proj_dims* dims, const void* x, const float* ln_w, const float* W, const float* dy,
                          const void* saved, void* scratch, float* dx, float* d_ln_w, float* d_ln_b, float* dW,
                          float* dbias, void* stream);

/* Route mask of the missing-modality protocol: replaces build_route_mask_from_presence
 * (MIMIC-IV/PhenoModel/Partial/Cross_Attention/routing_and_heads.py:10-64) / build_route_mask_from_modalities
 * (.../Partial/Cross_Attention/main.py:109-132).  hasL/hasN/hasI: fp32 [B] (1 = available; NULL = available for all);
 * route_mask: fp32 [B,10] in ROUTES order, 1 iff every modality the route needs is present.  drop_bits: bit r set
 * zeroes route r for the whole batch (training-time route dropout, MortModel/.../main.py:3027-3033). */
int mmr_route_mask_from_presence(const float* hasL, const float* hasN, const float* hasI, int B, int drop_bits,
                                 float* route_mask, void* stream);

/* Training tail: replaces torch.nn.utils.clip_grad_norm_ + grads_are_finite + torch.optim.AdamW.step + EMA.update
 * (MortModel/Paired_Cross_Attention/main.py:3143-3165, 2886-2890, 58-108) with three device-side stages that
 * never synchronise the host: (1) mmr_grad_sqnorm accumulates sum g^2 of any number of tensor tables into
 * state->sumsq; (2) mmr_opt_prepare turns it into total norm, clip coefficient min(1, max_norm/(norm+1e-6)), a skip
 * flag (non-finite norm: nothing is updated and the step counter does not advance) and the AdamW bias corrections
 * of step+1; (3) mmr_opt_apply updates p, exp_avg, exp_avg_sq (and the EMA shadow when has_ema) of a table. */
typedef struct mmr_opt_tensor {   /* device pointers, fp32, n elements each; ema may be NULL */
  float* p; const float* g; float* m; float* v; float* ema; int64_t n;
} mmr_opt_tensor;
typedef struct mmr_opt_hyper {
  double lr, beta1, beta2, eps, weight_decay;   /* torch.optim.AdamW arguments (amsgrad=False, maximize=False) */
  double max_norm;                              /* clip_grad_norm_ max_norm; <= 0 disables clipping */
  double ema_decay;                             /* EMA.decay */
  int32_t has_ema, reserved;
} mmr_opt_hyper;
typedef struct mmr_opt_state {    /* DEVICE resident, caller-owned, zero-initialised once; persists across steps */
  double sumsq, bc1, bc2_sqrt;
  float norm, clip;
  int32_t step, skip;
} mmr_opt_state;
int mmr_grad_sqnorm(const mmr_opt_tensor* host_tensors, int n_tensors, mmr_opt_state* state, void* stream);
int mmr_opt_prepare(const mmr_opt_hyper* hp, mmr_opt_state* state, void* stream);
int mmr_opt_apply(const mmr_opt_tensor* host_tensors, int n_tensors, const mmr_opt_hyper* hp,
                  const mmr_opt_state* state, void* stream);
/* EMA.update alone: ema = decay * ema + (1 - decay) * p for every table entry (g, m, v ignored). */
int mmr_ema_update(const mmr_opt_tensor* host_tensors, int n_tensors, double decay, void* stream);

/* ---- loss tail: what consumes logits / alpha / R right after capsule routing ----------
 * Replaces, without any host synchronisation (the reference's coerce_rc_to_report calls .item() twice per step),
 *   MORT   death_logit_from_logits2 + label smoothing + BCEWithLogitsLoss + route-entropy bonus + route-uniformity
 *          penalty on alpha                               MortModel/Paired_Cross_Attention/main.py:1753-1755, 3084-3126
 *   PHENO  coerce_rc_to_report + assert_routing_over_routes + BCEWithLogitsLoss(pos_weight) + batch-mean routing
 *          entropy / uniformity terms      PhenoModel/Paired_Cross_Attention/main.py:1472-1564, 261-276, 2755-2812
 * loss = base - ent + uni.  Only `base` carries gradient in the reference (alpha is returned detached,
 * routing_and_heads.py:363, and coerce_rc_to_report detaches R, main.py:1483), so `dlogits` = d loss / d logits is
 * the whole backward of this step; logits rewritten by _safe_tensor (NaN/Inf) get zero gradient like nan_to_num.
 * Lambdas of 0 disable a term (the epoch warm-up gates are host logic).  Everything is fp32 except rc_raw. */
typedef struct mmr_loss_state {   /* DEVICE resident, caller-owned, zero-initialised once */
  float loss, base, ent, uni;     /* ent / uni already multiplied by their lambdas */
  float err_routes, err_k;        /* max |sum_r R - 1|, max |sum_k R - 1| of the sanitised input (coerce case selection) */
  float max_route_sum_err;        /* max |sum_r rc_report - 1|: assert_routing_over_routes fails when > atol */
  int32_t info;                   /* 0 no routing coefficients; 1 already p(route|phenotype); 2 sums to one over labels
                                     (the reference raises TypeError there, main.py:1519); 3 forced normalisation */
  int32_t nonfinite_logits;       /* entries _safe_tensor rewrote */
  uint32_t ticket;                /* internal */
  int32_t reserved[2];
} mmr_loss_state;
typedef struct mmr_loss_args {
  int32_t variant;                /* MMR_VARIANT_MORT (logits [B,2], y [B]) / MMR_VARIANT_PHENO (logits, y [B,K]) */
  int32_t B, K;
  int32_t rc_dtype;               /* MMR_DTYPE_F32 / MMR_DTYPE_BF16 of rc_raw */
  const float* logits;
  const float* y;
  const float* pos_weight;        /* [K] or NULL (PHENO) */
  const float* prim_acts;         /* [B,10] or NULL: alpha (MORT regularisers) */
  const void* rc_raw;             /* [B,10,K] or NULL: R as returned by the capsule head (PHENO) */
  const float* route_mask;        /* [B,10] or NULL */
  float label_smoothing;          /* MORT */
  float route_entropy_lambda, route_uniform_lambda;
  float atol;                     /* coerce_rc_to_report / assert_routing_over_routes tolerance (reference: 1e-3) */
  float* dlogits;                 /* [B,K] out or NULL (evaluation) */
  float* rc_report;               /* [B,10,K] fp32 out or NULL */
  mmr_loss_state* state;
  void* scratch;                  /* mmr_loss_scratch_bytes(B) bytes, 8-byte aligned */
} mmr_loss_args;
size_t mmr_loss_scratch_bytes(int B);
int mmr_loss_fwd_bwd(const mmr_loss_args* args, void* stream);

/* Evaluation-time routing statistics: replaces the per-batch `.cpu()` copies and host sums of
 * evaluate_epoch (MortModel/Paired_Cross_Attention/main.py:1916-1933, 2013-2016).  Accumulates IN PLACE, on the device:
 *   sums[0][r][k] += sum_b rc_raw[b][r][k]      sums[1][r][k] += sum_b rc_report[b][r][k]   (rc_report may be NULL)
 *   sums[2][r][k] += sum_b rc_raw[b][r][k] * prim_acts[b][r]       sums[3*10*K + r] += sum_b prim_acts[b][r]
 * sums: fp32 [3*10*K + 10], zero-initialised once per split; count (may be NULL): device uint64 += B.  rc_raw fp32 or bf16. */
int mmr_routing_stats_accumulate(const void* rc_raw, int rc_dtype, const float* rc_report, const float* prim_acts, int B,
                                 int K, float* sums, unsigned long long* count, void* stream);

/* Unit-test hook for the GEMM engines: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) with bf16 (dtype 1)
 * or fp32 (dtype 0) operands, fp32 output.  trans=1 computes C[M,N] = A[Kr,M]^T * B[Kr,N]
 * (the weight-gradient form, reduction over rows). */
int mmr_debug_gemm(int engine, int dtype, int trans, int M, int N, int K, const void* A,
                   const void* B, const float* bias, float* C, void* stream);

/* Tuning hook: average device milliseconds of `iters` launches of the persistent tcgen05 GEMM
 * C[M,N] = A[M,K] * B[N,K]^T (bf16 operands) with epilogue op 0 (+bias -> bf16), 1 (+bias, ReLU, sign
 * bits), 2 (sign-bit mask) or 4 (fp32 out).  M is padded to 128 rows by the caller's buffers. */
int mmr_bench_gemm(int op, int M, int N, int K, const void* A, const void* B, const float* bias,
                   void* C, uint32_t* bits, int iters, float* ms_out, void* stream);

/* Instrumentation (bench / profiling only).  mmr_launch_count: kernels launched by this library
 * since load.  mmr_prof_enable(1) brackets every launch group with CUDA events on the launching
 * stream; mmr_prof_collect synchronises them and returns summed device milliseconds and group counts
 * for 8 classes: 0 tcgen05 gemm, 1 tcgen05 wgrad, 2 attention fwd, 3 attention bwd, 4 SIMT gemm,
 * 5 routing, 6 whole fusion fwd call, 7 whole fusion bwd call. */
long long mmr_launch_count(void);
int mmr_prof_enable(int on);
int mmr_prof_collect(double* ms_by_class, long long* n_by_class);

/* sizeof() of the public structs, in declaration order: mmr_fusion_dims, mmr_routing_dims, mmr_routing_params,
 * mmr_routing_grads, mmr_opt_tensor, mmr_opt_hyper, mmr_opt_state, mmr_loss_state, mmr_loss_args, mmr_proj_dims.  Lets a foreign-language binding assert that its
 * mirror of the structs matches this build (returns the number of entries written, at most n). */
int mmr_abi_struct_sizes(size_t* out, int n);

int mmr_version(void);
const char* mmr_last_error_string(void);

#ifdef __cplusplus
}
#endif
#endif /* MMR_B200_H */
