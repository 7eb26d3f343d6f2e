/*
 * scan_b200 C ABI -- the drop-in boundary of the B200-native condgraph middle head.
 *
 * Conventions (they mirror the native-op convention of the reference's
 * fcos_core/csrc, e.g. csrc/cuda/SigmoidFocalLoss_cuda.cu:103-188 and csrc/SigmoidFocalLoss.h:10-41,
 * minus the at::Tensor types):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns and pre-allocates every buffer (torch on the Python side) and passes the CUDA
 *     stream the work must be enqueued on (`stream` is a cudaStream_t cast to void*; NULL = legacy stream);
 *   - the callee never allocates device memory, never synchronises and never touches another stream;
 *   - return value: 0 on success, a negative SCAN_E* code otherwise (`scan_strerror`); the Python shim
 *     turns it into RuntimeError, like AT_ERROR / AT_ASSERTM in the reference;
 *   - dynamic counts (number of sampled nodes, DBSCAN points, ...) are produced in device-side int32
 *     records and read by the caller when it needs them;
 *   - "rows" layout: the pixels of all FPN levels of all images as one [R, C] row-major fp32 matrix in
 *     the reference's own flattening order -- level first, then image, then y, then x
 *     (`features[l].permute(0,2,3,1).reshape(-1,C)` of loss.py:440 concatenated over levels as in
 *     loss.py:289-294, condgraph.py:348-351).  Row g of level l, image n, pixel (y,x):
 *         g = row_off[l] + (n*H_l + y)*W_l + x,   row_off[l] = N * sum_{j<l} H_j*W_j.
 *   - every entry point is safe to call twice on the same saved tensors (the reference back-propagates
 *     the source graph twice, engine/trainer.py:299,343): inputs are never modified.
 *
 * All file:line citations are relative to /root/reference/fcos_core/.
 */
#ifndef SCAN_B200_H_
#define SCAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCAN_ABI_VERSION 1
#define SCAN_MAX_LEVELS 8
#define SCAN_MAX_CLASSES 16 /* K = used_num_classes <= 16 (one tcgen05 N=16 tile) */

enum {
  SCAN_OK = 0,
  SCAN_EINVAL = -1,   /* bad argument (shape, alignment, K > SCAN_MAX_CLASSES ...) */
  SCAN_ECUDA = -2,    /* a CUDA runtime / driver call failed; see scan_last_cuda_error() */
  SCAN_ENOTSUP = -3,  /* configuration not supported by this build */
  SCAN_ECAPACITY = -4 /* caller-provided workspace too small */
};

int scan_abi_version(void);
const char* scan_strerror(int code);
const char* scan_last_cuda_error(void);
/* bytes of dynamic shared memory / SM count etc. are queried lazily; this forces it (returns 0 / SCAN_ECUDA) */
int scan_init(int device);
/* Copy a small (<= 4 MB, multiple of 4 bytes) buffer to the device with a kernel.  `src` may be PINNED HOST memory
 * (device-accessible under unified addressing): per-call metadata such as the padded ground-truth boxes then bypasses the copy
 * engine, where it would wait behind a training loop's input prefetch.  The host buffer must stay alive until the stream
 * has passed this point (the Python shim keeps it referenced through a recorded event). */
int scan_upload_small(const void* src, void* dst, int64_t bytes, void* stream);

/* ---- geometry ------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_levels;                     /* <= SCAN_MAX_LEVELS */
  int32_t n_images;                     /* N */
  int32_t h[SCAN_MAX_LEVELS];           /* H_l */
  int32_t w[SCAN_MAX_LEVELS];           /* W_l */
  int32_t stride[SCAN_MAX_LEVELS];      /* FPN stride of the level (MODEL.FCOS.FPN_STRIDES) */
} scan_levels_t;

/* ---- layout: NCHW <-> rows (replaces features[l].permute(0,2,3,1).reshape(-1,C), loss.py:440) -- */
/* nchw_host: host array of n_levels device pointers, level l is [N, C, H_l, W_l] fp32 contiguous. */
int scan_pack_rows(const scan_levels_t* lv, const void* const* nchw_host, int32_t channels,
                   float* rows, void* stream);
/* rows [R, C] -> nchw (overwrite when accumulate == 0, += otherwise): the backward of scan_pack_rows */
int scan_unpack_rows(const scan_levels_t* lv, const float* rows, int32_t channels,
                     void* const* nchw_host, int32_t accumulate, void* stream);
/* scan_unpack_levels: the same from one NHWC-dense [N*H_l*W_l, C] matrix PER LEVEL (HOST array of device pointers; the
 * gradients cuDNN's backward-data hands back level by level), all levels in one launch. */
int scan_unpack_levels(const scan_levels_t* lv, const void* const* rows_levels_host, int32_t channels,
                       void* const* nchw_host, void* stream);

/* ---- f1: GroupNorm(32) + ReLU of the head_in towers on the rows layout (condgraph.py:68-119: the
 *      nn.GroupNorm(32, C) + nn.ReLU that follow every tower convolution) ---------
 * x_levels_host: host array of n_levels device pointers, level l = that level's convolution output as
 * NHWC-dense [N*H_l*W_l, 256] fp32 (a torch channels_last tensor).  y_rows [R,256] receives
 * relu(group_norm(x)) of all levels in the rows order; stats [n_levels*N*32*2] receives (mean, rstd).
 * Backward: dy_levels_host like x_levels_host; dx_rows [R,256] in rows order; dgamma, dbeta [256].
 * Workspace: scan_gn_workspace_bytes(lv) bytes.  Deterministic (per-block partials combined in fp64). */
int64_t scan_gn_workspace_bytes(const scan_levels_t* lv);
/* conv_bias (nullable, [256]): the bias of the preceding convolution, added on the fly so that the convolution
 * itself runs bias-free; the backward then also returns d_conv_bias (column sums of dx) instead of a
 * separate full-tensor reduction.  conv_bias and d_conv_bias must both be given or both be NULL.
 * The backward does not read y: it recomputes the ReLU mask [x * a + c > 0] from x with the forward's own
 * operations (a = rstd * gamma, c = beta - mean * a, one fused multiply-add), which is why it takes beta. */
int scan_gn_relu_fwd(const scan_levels_t* lv, const void* const* x_levels_host, const float* conv_bias,
                     const float* gamma, const float* beta, float eps, float* y_rows, float* stats,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* the apply pass alone (y = relu(gn(x + conv_bias))) with statistics from scan_conv3x3_rows_gn */
int scan_gn_relu_apply(const scan_levels_t* lv, const void* const* x_levels_host, const float* conv_bias,
                       const float* gamma, const float* beta, const float* stats, float* y_rows, void* stream);
int scan_gn_relu_bwd(const scan_levels_t* lv, const void* const* x_levels_host,
                     const void* const* dy_levels_host, const float* conv_bias, const float* gamma,
                     const float* beta, const float* stats, float* dx_rows, float* dgamma, float* dbeta,
                     float* d_conv_bias, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- f1: head_out epilogue y = relu(u + v + bias) (condgraph.py:379-384 with the concat removed:
 *      u = conv3x3(features, W[:, :256]), v = conv3x3(act maps, W[:, 256:])) --------------------------
 * u_levels_host / v_levels_host (v nullable): per-level NHWC-dense [N*H_l*W_l, 256]; y_rows [R,256].
 * Backward: d_rows [R,256] = dy * [y > 0] (the gradient of both u and v), d_bias [256] (nullable;
 * needs a workspace of scan_gn_workspace_bytes(lv)). */
int scan_add_relu_fwd(const scan_levels_t* lv, const void* const* u_levels_host,
                      const void* const* v_levels_host, const float* bias, float* y_rows, void* stream);
int scan_add_relu_bwd(const scan_levels_t* lv, const void* const* dy_levels_host, const float* y_rows,
                      float* d_rows, float* d_bias, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K1a: FCOS ground-truth assignment (loss.py:262-343, PrototypeComputation.prepare_targets +
 *      compute_targets_for_locations; locations of condgraph.py:631-655 are computed from the index) --
 * boxes      [N, g_max, 4] fp32 xyxy, box_labels [N, g_max] int64, box_count [N] int32 (1..g_max)
 * labels_out [R] int64, rows layout.  Bit-exact with the reference (fp32, no FMA contraction). */
int scan_fcos_assign(const scan_levels_t* lv, const float* boxes, const int64_t* box_labels,
                     const int32_t* box_count, int32_t g_max, int64_t* labels_out, void* stream);
/* The same assignment for FCOSLossComputation (loss.py:40-126), which also needs the regression targets:
 * reg_targets_out [R, 4] = (l, t, r, b) of the chosen box (box 0 where no box matched, like the reference's argmin). */
int scan_fcos_assign_reg(const scan_levels_t* lv, const float* boxes, const int64_t* box_labels,
                         const int32_t* box_count, int32_t g_max, int64_t* labels_out, float* reg_targets_out,
                         void* stream);

/* ---- f2: FCOSLossComputation.__call__ fused (loss.py:168-230; layers/iou_loss.py:5-38; SigmoidFocalLoss_cuda.cu:36-98) ----
 * Reads the FCOS head's NCHW maps in place (HOST arrays of per-level device pointers: cls [N,C,H,W] logits, reg [N,4,H,W],
 * ctr [N,1,H,W] logits) with labels [R] / reg_targets [R,4] from scan_fcos_assign_reg.
 * losses3 = (cls_loss, reg_loss, centerness_loss) device scalars; sums6 (fp64: focal sum, #pos, sum w, sum iou*w, sum iou,
 * sum bce) is kept for the backward; partials: scan_fcos_loss_num_partials() doubles of scratch. */
int32_t scan_fcos_loss_num_partials(void);
int scan_fcos_loss_fwd(const scan_levels_t* lv, const void* const* cls_host, const void* const* reg_host, const void* const* ctr_host,
                       const int64_t* labels, const float* reg_targets, int32_t num_classes, float gamma, float alpha,
                       double* partials, double* sums6, float* losses3, void* stream);
/* d_losses3: device vector of the three upstream gradients; writes d_cls / d_reg / d_ctr (same shapes as the maps) */
int scan_fcos_loss_bwd(const scan_levels_t* lv, const void* const* cls_host, const void* const* reg_host, const void* const* ctr_host,
                       const int64_t* labels, const float* reg_targets, int32_t num_classes, float gamma, float alpha,
                       const double* sums6, const float* d_losses3, void* const* d_cls_host, void* const* d_reg_host,
                       void* const* d_ctr_host, void* stream);


/* ---- K1b: node sampling (loss.py:430-458 source branch; loss.py:497-516 target branch) ---------
 * mode 0 (source): positive <=> labels[g] > 0, node label = labels[g]; per level all negatives when
 *                  n_pos > n_neg, else the floor(linspace(0, n_neg-2, n_pos)) rows of the negative list;
 *                  with_bg == 0 drops the negatives (PROTO_WITH_BG False).
 * mode 1 (target): positive <=> pos_mask[g] != 0, node label = plabel[g]; levels without a positive
 *                  contribute nothing; n_pos negatives by floor(linspace(0, n_neg-2, n_pos)) (negative
 *                  indices wrap like Python indexing; n_neg == 0 sets meta.error = 1: the reference
 *                  raises IndexError there).
 * Node order: [neg(level 0..), pos(level 0..)].  Outputs: node_rows int32 [cap], node_labels int64 [cap],
 * meta (device): see scan_sample_meta_t.  workspace: scan_sample_workspace_bytes(R). */
typedef struct {
  int32_t n_nodes;                      /* M */
  int32_t n_neg_nodes;                  /* nodes [0, n_neg_nodes) are negatives */
  int32_t error;                        /* 0 ok, 1 = no negative row left at a level with positives, 2 = cap too small */
  int32_t reserved;
  int32_t n_pos[SCAN_MAX_LEVELS];
  int32_t n_neg[SCAN_MAX_LEVELS];
  int32_t n_neg_sel[SCAN_MAX_LEVELS];
  int32_t neg_off[SCAN_MAX_LEVELS];     /* first node index of the level's negatives */
  int32_t pos_off[SCAN_MAX_LEVELS];     /* first node index of the level's positives */
} scan_sample_meta_t;

int64_t scan_sample_workspace_bytes(int64_t n_rows);
int scan_sample_nodes(const scan_levels_t* lv, int32_t mode, int32_t with_bg, const int64_t* labels,
                      const uint8_t* pos_mask, const int64_t* plabel, int32_t* node_rows,
                      int64_t* node_labels, int32_t cap, scan_sample_meta_t* meta, void* workspace,
                      int64_t workspace_bytes, void* stream);

/* ---- K1c: gather / scatter of node rows ------------------------------------------------------ */
/* out[i, :] = rows[node_rows[i], :], i < n_nodes (C % 4 == 0) */
int scan_gather_rows(const float* rows, const int32_t* node_rows, int32_t n_nodes, int32_t channels,
                     float* out, void* stream);
/* d_rows[node_rows[i], :] += d_nodes[i, :] (repeated indices allowed) */
int scan_scatter_add_rows(const float* d_nodes, const int32_t* node_rows, int32_t n_nodes,
                          int32_t channels, float* d_rows, void* stream);

/* ---- K4b: conditional 1x1 convolution + activation + focal loss (condgraph.py:619-629 dynamic_conv,
 *      :338-370 get_act_loss; layers/sigmoid_focal_loss_wbg.py:7-64 FocalLoss, :148-177 BCEFocalLoss) --
 * rows [R, 256] fp32, weight [K, 256], bias [K] or NULL (COND_WITH_BIAS).
 * act_mode 0: softmax over K (softmaxFL / no loss), 1: sigmoid (sigmoidFL).
 * act_nchw_host: host array of n_levels device pointers, level l is [N, K, H_l, W_l]: the activation maps
 *                in the reference's own layout.
 * labels [R] int64 or NULL; when given, loss_partials[i] (fp64, i < scan_condconv_num_partials()) receives
 * per-CTA partial sums of the UN-normalised focal loss (caller divides by R for softmaxFL and by 2R for
 * sigmoidFL and multiplies by ACT_LOSS_WEIGHT); flags[i] (int32, same length as loss_partials) is non-zero
 * when CTA i clamped a p_t < 1e-15.
 * Arithmetic: impl 0 = tcgen05.mma kind::tf32, 3xTF32 error compensation, activation operands in tensor memory,
 * fp32 accumulation in TMEM (the product); impl 1 = fp32 FFMA verification kernel; impl 2 = as 0 with the operands in
 * shared memory.  Tolerance of impl 0 / 2 against the fp32 reference: rtol 1e-3 (measured 3e-5). */
int32_t scan_condconv_num_partials(void);
int scan_condconv_fwd(const scan_levels_t* lv, const float* rows, const float* weight, const float* bias,
                      int32_t num_classes, int32_t act_mode, void* const* act_nchw_host,
                      const int64_t* labels, double* loss_partials, int32_t* flags, int32_t impl,
                      void* stream);
/* Backward.  d_act_nchw_host: upstream gradient w.r.t. the activation maps (NULL entries = zero),
 * loss_scale: ACT_LOSS_WEIGHT / normaliser (0 when there is no act loss); d_loss: device scalar
 * d(total)/d(act_loss) or NULL (= 1), read on the device so that backward needs no host sync.
 * Outputs: d_rows [R,256] (overwritten), d_weight [K,256] and d_bias [K] (overwritten; d_bias may be NULL).
 * workspace: scan_condconv_bwd_workspace_bytes(lv, num_classes) (d_weight partials + the [R, num_classes] d_logit matrix). */
int64_t scan_condconv_bwd_workspace_bytes(const scan_levels_t* lv, int32_t num_classes);
int scan_condconv_bwd(const scan_levels_t* lv, const float* rows, const float* weight,
                      int32_t num_classes, int32_t act_mode, const void* const* act_nchw_host,
                      const void* const* d_act_nchw_host, const int64_t* labels, float loss_scale,
                      const float* d_loss, float* d_rows, float* d_weight, float* d_bias, void* workspace,
                      int64_t workspace_bytes, void* stream);
/* the same with accumulate_rows != 0: d_rows already holds the gradient of a LATER consumer of the same rows (head_out's data
 * gradient) and this call adds its own in place — the sum autograd would otherwise form with a separate full-size add. */
int scan_condconv_bwd2(const scan_levels_t* lv, const float* rows, const float* weight, int32_t num_classes, int32_t act_mode,
                       const void* const* act_nchw_host, const void* const* d_act_nchw_host, const int64_t* labels, float loss_scale,
                       const float* d_loss, float* d_rows, int32_t accumulate_rows, float* d_weight, float* d_bias, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* ---- K4a: paradigm manifestation, RNN variant (condgraph.py:313-336 get_conded_weight with USE_RNN:
 *      nn.RNN(I->H, 2 layers, tanh) over the P paradigm slots of the K classes, then the (P x 1)
 *      convolution cond_nx1) --------------------------------------------------------------------------
 * proto [K, I, P] (the `prototype` buffer), RNN parameters in torch's nn.RNN layout (weight_ih_l0 [H,I],
 * weight_hh_l0 [H,H], biases [H], layer 1 [H,H]), wc = cond_nx1.weight [O, H, P, 1], bc [O].
 * kernel_out [K, O]; `saved` (scan_manifest_rnn_saved_floats floats) keeps the activations for backward.
 * K <= 16, P <= 16, I, H and O multiples of 4.  Backward returns the gradients of all ten parameters
 * (the prototype buffer takes no gradient in the reference). */
int64_t scan_manifest_rnn_saved_floats(int32_t K, int32_t P, int32_t I, int32_t H);
int64_t scan_manifest_rnn_workspace_bytes(int32_t K, int32_t P, int32_t H, int32_t O);
int scan_manifest_rnn_fwd(const float* proto, int32_t K, int32_t P, int32_t I, int32_t H, int32_t O,
                          const float* w_ih0, const float* w_hh0, const float* b_ih0, const float* b_hh0,
                          const float* w_ih1, const float* w_hh1, const float* b_ih1, const float* b_hh1,
                          const float* wc, const float* bc, float* kernel_out, float* saved, void* stream);
int scan_manifest_rnn_bwd(const float* d_kernel, int32_t K, int32_t P, int32_t I, int32_t H, int32_t O,
                          const float* w_hh0, const float* w_ih1, const float* w_hh1, const float* wc,
                          const float* saved, float* d_w_ih0, float* d_w_hh0, float* d_b_ih0, float* d_b_hh0,
                          float* d_w_ih1, float* d_w_hh1, float* d_b_ih1, float* d_b_hh1, float* d_wc,
                          float* d_bc, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K3a: graph aggregation, global variant (layers/transformer.py:5-90 dot_attention inside
 *      MultiHeadAttention, called at condgraph.py:390-393) ----------------------------------------
 * q,k,v [M,256] are the three linear projections.  The reference's .view(4,-1,64) makes 4 independent
 * chunks of M sub-tokens of 64 dims: chunk b owns sub-token rows [b*M,(b+1)*M) of the
 * row-major [4M,64] reinterpretation.  ctx [M,256] (same reinterpretation), lse [4M] (log-sum-exp of the
 * scaled scores, saved for backward).  scale = 0.25.  dropout_p > 0 applies a counter-based Bernoulli
 * mask keyed by (seed, chunk, i, j) to the probabilities (train-mode nn.Dropout of transformer.py:31).
 * Forward: with a workspace of scan_attn_workspace_bytes(m) the tcgen05 kernel runs (3xTF32, fp32 accumulate in TMEM);
 * workspace == NULL selects the fp32 FFMA kernel (verification path). */
int64_t scan_attn_workspace_bytes(int32_t m);
int scan_attn_fwd(const float* q, const float* k, const float* v, int32_t m, float scale,
                  float dropout_p, uint64_t seed, float* ctx, float* lse, void* workspace,
                  int64_t workspace_bytes, void* stream);
/* Backward: delta_ws [4M] floats of scratch.  With a workspace of scan_attn_bwd_workspace_bytes(m) the two tcgen05
 * kernels run (dQ per 128-query CTA, dK/dV per 128-key CTA, no atomics); workspace == NULL selects the FFMA kernel. */
int64_t scan_attn_bwd_workspace_bytes(int32_t m);
int scan_attn_bwd(const float* q, const float* k, const float* v, const float* ctx, const float* lse,
                  const float* d_ctx, int32_t m, float scale, float dropout_p, uint64_t seed,
                  float* dq, float* dk, float* dv, float* delta_ws, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* ---- K3a': the dense layers around the attention core and the node classifier, tcgen05 3xTF32 GEMMs (csrc/gemm.cu).
 *      Replaces nn.Linear / nn.LayerNorm / nn.Dropout / F.cross_entropy of layers/transformer.py:43-49, 61-63, 84-88 and
 *      condgraph.py:186-188, 400-402 (cuBLAS fp32 SIMT + ATen kernels in torch).  Channel width fixed at 256.
 * scan_graph_workspace_bytes(m): workspace size accepted by every *_bwd entry point below for m nodes. */
int64_t scan_graph_workspace_bytes(int32_t m);
/* qkv [3, M, 256] (q | k | v, each a contiguous [M,256] matrix) = x [M,256] . w_qkv [768,256]^T + b_qkv [768] */
int scan_qkv_fwd(const float* x, const float* w_qkv, const float* b_qkv, int32_t m, float* qkv, void* stream);
/* d_x (+)= d_qkv . w_qkv (accumulate_dx != 0: added to the residual-branch gradient already in d_x);
 * d_w_qkv [768,256] = d_qkv^T . x;  d_b_qkv [768] = column sums (deterministic). */
int scan_qkv_bwd(const float* d_qkv, const float* x, const float* w_qkv, int32_t m, int32_t accumulate_dx, float* d_x,
                 float* d_w_qkv, float* d_b_qkv, void* workspace, int64_t workspace_bytes, void* stream);
/* y = LayerNorm(x + dropout(ctx . w_f^T + b_f)) * gamma + beta, fused in the GEMM epilogue (one TMEM lane = one node);
 * the dropout mask is a counter hash of (seed, row, column); xhat [M,256] and rstd [M] are saved for the backward. */
int scan_attn_out_ln_fwd(const float* ctx, const float* w_f, const float* b_f, const float* x, const float* gamma,
                         const float* beta, int32_t m, float eps, float drop_p, uint64_t seed, float* y, float* xhat,
                         float* rstd, void* stream);
/* d_y -> d_x (residual branch), d_ctx, d_w_f [256,256], d_b_f [256], d_gamma_beta [512] = d_gamma | d_beta */
int scan_attn_out_ln_bwd(const float* d_y, const float* xhat, const float* rstd, const float* gamma, const float* ctx,
                         const float* w_f, int32_t m, float drop_p, uint64_t seed, float* d_x, float* d_ctx, float* d_w_f,
                         float* d_b_f, float* d_gamma_beta, void* workspace, int64_t workspace_bytes, void* stream);
/* loss (device scalar) = loss_weight * mean_m CE(relu(nodes . w1^T + b1) . w2^T + b2, labels - label_shift);
 * hidden [M,512] and dlogits [M,16] (= softmax - onehot, columns >= K zero) are saved for the backward.
 * workspace >= 8 * ceil(M / 128) + 256 bytes. */
int scan_node_cls_fwd(const float* nodes, const float* w1, const float* b1, const float* w2, const float* b2,
                      const int64_t* labels, int32_t m, int32_t hidden_dim, int32_t num_classes, int32_t label_shift,
                      float loss_weight, float* hidden, float* dlogits, float* loss, void* workspace, int64_t workspace_bytes,
                      void* stream);
/* d_loss: device scalar d(total)/d(loss).  Outputs d_nodes [M,256], d_w1 [512,256], d_b1 [512], d_w2 [K,512], d_b2 [K]. */
int scan_node_cls_bwd(const float* dlogits, const float* hidden, const float* nodes, const float* w1, const float* w2,
                      int32_t m, int32_t hidden_dim, int32_t num_classes, float loss_weight, const float* d_loss,
                      float* d_nodes, float* d_w1, float* d_b1, float* d_w2, float* d_b2, void* workspace,
                      int64_t workspace_bytes, void* stream);
/* backward of the per-class node means (condgraph.py:395-398): d_nodes[m,:] = d_mean[label[m] - shift,:] / count */
int scan_class_mean_bwd(const float* d_mean, const float* packed_sums, const int64_t* labels, int32_t m, int32_t channels,
                        int32_t label_shift, float* d_nodes, void* stream);

/* ---- K3b: per-class prototype sums + paradigm EMA (condgraph.py:395-398 class means;
 *      :558-617 update_prototype / _nx1 / _nx1_rnn) -------------------------------
 * scan_class_sums: sums[c, :] = sum of nodes with label c + label_shift... , sums[K*C + c] = count
 *   nodes [M, C], labels [M] int64 (class c <=> labels == c + label_shift), out: packed [K, C+1] fp32
 *   (row c = sum over nodes, last column = count) -- the buffer the multi-GPU all-reduce operates on. */
int scan_class_sums(const float* nodes, const int64_t* labels, int32_t m, int32_t channels,
                    int32_t num_classes, int32_t label_shift, float* packed_sums, void* stream);
/* scan_proto_update: batch[c] = sums[c]/count[c] (0 where count == 0), exist = (sum_j batch[c][j] != 0),
 *   then EMA into prototype [K, C, P] (P == 1: [K, C]) slot `slot` with momentum = cosine(old, batch)
 *   (cosine_on) or `momentum`; shift != 0 first moves slots 1..P-1 down by one for ALL classes
 *   (counter == PROTO_ITER case of update_prototype_nx1_rnn).  proto_batch_out [K, C] receives batch. */
int scan_proto_update(const float* packed_sums, int32_t num_classes, int32_t channels, int32_t proto_iter,
                      int32_t slot, int32_t shift, int32_t cosine_on, float momentum, float* prototype,
                      float* proto_batch_out, void* stream);

/* ---- a6: building blocks of the per-class GCN (GLOBAL_GCN = False; condgraph.py:262-302, 404-414) -------------------
 * All on the tcgen05 3xTF32 GEMM of csrc/gemm.cu; pitches in floats, multiples of 4; pointers 16-byte aligned.
 * scan_gemm_nt: c [m,n] (+)= act(a [m,k] . b [n,k]^T + bias[n]); k > 512 is split and reduced in a fixed order
 *   (workspace of scan_gemm_nt_workspace_bytes(m, n, k) bytes, may be NULL when k <= 512). */
int64_t scan_gemm_nt_workspace_bytes(int32_t m, int32_t n, int32_t k);
int scan_gemm_nt(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t m, int32_t n, int32_t k, const float* bias,
                 int32_t relu, int32_t accumulate, float* c, int64_t ldc, void* workspace, int64_t workspace_bytes, void* stream);
/* dst [n_cols, ld_dst] = src [n_rows, n_cols]^T, zero-padded to ld_dst columns */
int scan_transpose(const float* src, int32_t n_rows, int32_t n_cols, int64_t ld_src, float* dst, int64_t ld_dst, void* stream);
/* d_w [n_out,n_in] (+)= dz [m,n_out]^T . x [m,n_in]; d_b [n_out] (+)= column sums of dz (d_b may be NULL) */
int64_t scan_linear_wgrad_workspace_bytes(int32_t m, int32_t n_out, int32_t n_in);
int scan_linear_wgrad(const float* dz, const float* x, int32_t m, int32_t n_out, int32_t n_in, int32_t accumulate, float* d_w,
                      float* d_b, void* workspace, int64_t workspace_bytes, void* stream);
/* in-place softmax over the first n_cols entries of each of n_rows rows (pitch ld); the pad columns are zeroed (get_edge :284-302) */
int scan_rows_softmax(float* x, int32_t n_rows, int32_t n_cols, int64_t ld, void* stream);
/* y [m,256] = x / max(|x|_2, eps) row-wise (sim_matrix, condgraph.py:35-43) */
int scan_rows_l2normalize(const float* x, int32_t m, float eps, float* y, void* stream);
/* GCN output activation over [m,256] rows: mode 0 NO, 1 relu, 2 sigmoid, 3 tanh, 4 softmax(dim=-1) (condgraph.py:274-281);
 * act_out = act(z) (saved for the backward), y = act_out + shortcut (shortcut may be NULL; y may alias act_out when it is) */
int scan_gcn_act_fwd(const float* z, const float* shortcut, int32_t m, int32_t mode, float* act_out, float* y, void* stream);
int scan_gcn_act_bwd(const float* act_out, const float* dy, int32_t m, int32_t mode, float* dz, void* stream);

/* ---- K4a': manifestation without the RNN (condgraph.py:320-334): tiny-batch dense layers over the K <= 16 paradigm rows ----
 * y [K, O] = act(x [K, I] . w [O, I]^T + b); relu != 0 applies ReLU.  I % 4 == 0, K * I * 4 bytes <= 200 KB. */
int scan_rows_linear_fwd(const float* x, const float* w, const float* b, int32_t k, int32_t in_dim, int32_t out_dim, int32_t relu,
                         float* y, void* stream);
/* y = the forward output (read only when relu != 0); d_b and d_x may be NULL */
int scan_rows_linear_bwd(const float* x, const float* w, const float* dy, const float* y, int32_t k, int32_t in_dim, int32_t out_dim,
                         int32_t relu, float* d_w, float* d_b, float* d_x, void* stream);
/* y = relu(group_norm(x [K, C], groups)) per row (F.group_norm on a 2-D tensor, condgraph.py:326-327); stats [K, groups, 2] */
int scan_rows_gn_relu_fwd(const float* x, const float* gamma, const float* beta, int32_t k, int32_t channels, int32_t groups, float eps,
                          float* y, float* stats, void* stream);
int scan_rows_gn_relu_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* stats, int32_t k,
                          int32_t channels, int32_t groups, float* d_x, float* d_gamma, float* d_beta, void* stream);

/* ---- f1: the 3x3 convolutions of the GRAPHHead towers (condgraph.py:68-119; head_in :549, head_out :383) as a tcgen05 implicit
 *      GEMM on the rows layout, all levels and images in one launch ---------------------------------------------------------
 * scan_conv3x3_pack_weights: w[co][ci][ky][kx] (element strides s_*) -> packed[9][rows_pad][cols_pad] (rows padded to 256, columns
 * to 32, zero filled; scan_conv3x3_packed_floats(rows, cols) floats).  transpose = 0: rows = co, cols = ci (forward);
 * transpose = 1: rows = ci, cols = co and the taps rotated by 180 degrees (data gradient = the same convolution of dY).
 * packed_lo (may be NULL) receives the 3xTF32 residual plane.  scan_tf32_residual: lo = rna_tf32(x - trunc_tf32(x)). */
int64_t scan_conv3x3_packed_floats(int32_t rows, int32_t cols);
int scan_conv3x3_pack_weights(const float* w, int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx, int32_t cout, int32_t cin,
                              int32_t transpose, float* packed_hi, float* packed_lo, void* stream);
int scan_tf32_residual(const float* x, int64_t n, float* lo, void* stream);
/* y_rows [R, ldo] (first n_out columns) = act(conv3x3(x_rows [R, cin], padding 1) + bias + addend); cin % 32 == 0.  x_lo and
 * packed_lo both non-NULL: 3xTF32 (fp32-accurate, the parity mode); both NULL: single-pass TF32 (what cuDNN runs under torch's
 * default allow_tf32).  cta_group 2 = CTA pairs sharing the weight tile (the fast path), 1 = single CTAs. */
int scan_conv3x3_rows(const scan_levels_t* lv, const float* x_rows, const float* x_lo, int32_t cin, const float* packed,
                      const float* packed_lo, int32_t n_out, const float* bias, const float* addend, int32_t relu, float* y_rows,
                      int32_t ldo, int32_t cta_group, void* stream);
/* the tower convolution in front of a GroupNorm(32) (condgraph.py:100-107): y_rows [R, 256] = conv3x3(x_rows), bias-free, plus the
 * GroupNorm statistics of (y + gn_bias) from the kernel's epilogue (per-warp partial sums, combined per (level, image, group) in
 * fp64 in a fixed order): stats [L * N * 32][mean, 1 / sqrt(var + eps)] -- the array scan_gn_relu_fwd's statistics pass would
 * produce, without re-reading the tensor.  scan_gn_relu_apply then normalises with them. */
int64_t scan_conv3x3_gn_workspace_bytes(const scan_levels_t* lv);
int scan_conv3x3_rows_gn(const scan_levels_t* lv, const float* x_rows, const float* x_lo, int32_t cin, const float* packed,
                         const float* packed_lo, const float* gn_bias, float eps, float* y_rows, float* stats, int32_t cta_group,
                         void* workspace, int64_t workspace_bytes, void* stream);
/* the same with an optional SECOND input tensor x2_rows [R, cin2] whose channels follow x_rows' in the weight columns (the
 * concatenation [features | activation maps] of head_out, condgraph.py:379-384, and of the CKA discriminator's class-conditional
 * maps, fcos_head_discriminator_con.py:104-105, without materialising it), and an optional mask [R, ldo]: out = mask > 0 ? out : 0
 * (data gradient through the ReLU of a saved forward output). */
int scan_conv3x3_rows2(const scan_levels_t* lv, const float* x_rows, const float* x_lo, int32_t cin, const float* x2_rows,
                       const float* x2_lo, int32_t cin2, const float* packed, const float* packed_lo, int32_t n_out, const float* bias,
                       const float* addend, const float* mask, int32_t relu, float* y_rows, int32_t ldo, int32_t cta_group, void* stream);

/* y_rows [R, ldo] (first n_out columns) = act(x_rows [R, cin] . w^T + bias) on the same kernel with one tap (no shifted reads):
 * w [256-padded rows, cin] row-major with zero rows beyond n_out; w_lo non-NULL (with x_lo): 3xTF32. */
int scan_conv1x1_rows(const scan_levels_t* lv, const float* x_rows, const float* x_lo, int32_t cin, const float* w, const float* w_lo,
                      int32_t n_out, const float* bias, int32_t relu, float* y_rows, int32_t ldo, int32_t cta_group, void* stream);
/* weight gradient of the same convolution: d_w[co][ci][ky][kx] (element strides s_*) = sum_p dy_rows[p, co] * x_rows[p + off, ci];
 * cin and cout multiples of 256; both operands are read in place (MN-major tcgen05 operands, no transposed copy); x_lo / dy_lo
 * both non-NULL: 3xTF32.  Deterministic (per-CTA-pair partial tiles summed in a fixed order); workspace from the _bytes query. */
int64_t scan_conv3x3_wgrad_workspace_bytes(const scan_levels_t* lv, int32_t cin, int32_t cout, int32_t precise);
int scan_conv3x3_wgrad(const scan_levels_t* lv, const float* x_rows, const float* x_lo, int32_t cin, const float* dy_rows,
                       const float* dy_lo, int32_t cout, float* d_w, int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* the class-map columns of head_out's weight gradient: d_w[co][k][ky][kx] (element strides; co < 256, k < K, 9 K <= 128) =
 * sum_p dy_rows[p, co] * maps32[p + off, k].  The maps are spread to [R, 128] (one column per (class, tap)) and reduced against
 * dy_rows by ONE MN-major tcgen05 GEMM over the pixels.  dy_lo non-NULL: 3xTF32.  Deterministic. */
int64_t scan_thin_wgrad_workspace_bytes(const scan_levels_t* lv, int32_t precise);
int scan_thin_wgrad(const scan_levels_t* lv, const float* maps32, const float* dy_rows, const float* dy_lo, int32_t k, float* d_w,
                    int64_t s_co, int64_t s_k, int64_t s_ky, int64_t s_kx, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- f3: thin kernels of the CKA discriminator FCOSDiscriminator_con (modeling/discriminator/fcos_head_discriminator_con.py:88-127,
 *      layer.py:6-24); its convolutions are scan_conv3x3_rows2 / scan_conv3x3_wgrad ---------------------------------------------
 * scan_thin_pack: channels [c0, c0 + k) of per-level NCHW maps [N, k_total, H_l, W_l] (HOST array of device pointers) -> rows
 * [R, ld] columns [0, k), the remaining columns zero.  scan_thin_unpack: the inverse, times `scale` (the other channels of the
 * NCHW tensors are left untouched).  scan_scale: y = scale * x (gradient reversal).  scan_colsum: out[c] = sum_r x[r, c]. */
int scan_thin_pack(const scan_levels_t* lv, const void* const* nchw_host, int32_t k_total, int32_t c0, int32_t k, float* rows, int32_t ld,
                   void* stream);
int scan_thin_unpack(const scan_levels_t* lv, const float* rows, int32_t ld, int32_t k_total, int32_t c0, int32_t k, float scale,
                     void* const* nchw_host, void* stream);
int scan_scale(const float* x, int64_t n, float scale, float* y, void* stream);
/* thin data gradient of a 3x3 convolution with K <= 14 input channels: d [R, ldd] = dY . W[:, k, tap]^T (one scan_gemm_nt of the
 * pixel rows against the [K * 9, C] weight slice, column k * 9 + tap); out32[p, k] = sum_tap d[p - off(tap), k * 9 + tap]. */
int scan_thin_gather(const scan_levels_t* lv, const float* d, int32_t ldd, int32_t k, float* out32, void* stream);
int64_t scan_colsum_workspace_bytes(int64_t n_rows, int32_t n_cols);
int scan_colsum(const float* x, int64_t n_rows, int32_t n_cols, int32_t ld, float* out, void* workspace, int64_t workspace_bytes,
                void* stream);
/* class-weighted BCE with logits (:113-121): n_cls > 1: loss = sum_c [sum_p w_c bce(x_c, t) / sum_p w_c] / n_cls with w = the class
 * maps (detached); n_cls == 1: the plain mean.  inv [16] receives the per-class gradient factors; scan_cka_bce_bwd writes
 * d_logits = d_loss * (sigmoid(x) - t) * w * inv[c] to dl32 [R, 32] (zero padded) and, if non-NULL, to dl_wide [R, ld_wide]. */
int64_t scan_cka_bce_workspace_bytes(void);
int scan_cka_bce_fwd(const float* logits, int32_t ldl, const float* weights, int32_t ldw, int64_t n_rows, int32_t n_cls, float target,
                     float* loss, float* inv, void* workspace, int64_t workspace_bytes, void* stream);
int scan_cka_bce_bwd(const float* logits, int32_t ldl, const float* weights, int32_t ldw, int64_t n_rows, int32_t n_cls, float target,
                     const float* inv, const float* d_loss, float* dl32, float* dl_wide, int32_t ld_wide, void* stream);

/* ---- f4: FCOS post-processor (modeling/rpn/fcos/inference.py:54-194; boxlist_nms structures/boxlist_ops.py:9-31; IoU and
 *      suppression rule csrc/cuda/nms.cu:13-67) on the probability maps of scan_ensemble_levels ----------------------------
 * prob [N,C,H_l,W_l] probabilities, reg [N,4,H_l,W_l], ctr [N,1,H_l,W_l] logits (HOST arrays of device pointers), image_hw [N,2]
 * int32 (h, w) on the device.  Per (image, level): candidates prob > pre_nms_thresh in (location, class) order, top
 * pre_nms_top_n (<= 1024) by prob * sigmoid(ctr), decode + clip_to_image + min-size filter, score = sqrt(.); per image and class
 * greedy NMS (IoU > nms_thresh, +1 pixel convention), then keep score >= the post_top_n-th largest.
 * Outputs (capacity 8192 per image): out_box [N,8192,4], out_score [N,8192], out_label [N,8192] int32 (1-based), out_count [N]
 * int32, ordered like the reference's result (class ascending, candidate order inside a class).  No host synchronisation. */
int64_t scan_postprocess_workspace_bytes(int32_t n_images, int32_t n_levels);
int scan_postprocess(const scan_levels_t* lv, const void* const* prob_host, const void* const* reg_host, const void* const* ctr_host,
                     int32_t num_classes, const int32_t* image_hw, float pre_nms_thresh, int32_t pre_nms_top_n, float nms_thresh,
                     int32_t post_top_n, float min_size, float* out_box, float* out_score, int32_t* out_label, int32_t* out_count,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a14: transfer (graph-matching) losses of the target branch (condgraph.py:457-498, sim_matrix :35-43) ----------
 * prototype: the module's paradigm buffer [K, 256, P] (P == 1: [K, 256]); sr_proto = prototype.mean(-1) is taken in-kernel.
 * NODES: loss = KLDivLoss(reduction='mean')(softmax(nodes).log(), softmax(sr_proto[labels])); diff [M,256] = softmax(nodes) -
 * target is saved and IS the gradient up to the scalar d_loss / (M * 256).  partials: scan_transfer_nodes_num_partials() doubles. */
int32_t scan_transfer_nodes_num_partials(void);
int scan_transfer_nodes_fwd(const float* nodes, const int64_t* labels, const float* prototype, int32_t proto_iter, int32_t m,
                            int32_t num_classes, float* diff, double* partials, float* loss, void* stream);
int scan_transfer_nodes_bwd(const float* diff, const float* d_loss, int32_t m, float* d_nodes, void* stream);
/* flags bit 0 PROTOTYPE (KL over the classes present in the target batch), bit 1 ADJ, bit 2 ADJ_COMPLETE (1 - cosine of the
 * flattened class-affinity matrices).  "present" = tg_proto row sum != 0, kept on the device (the reference's boolean-mask
 * indexing synchronises).  losses4 = [PROTOTYPE, ADJ, ADJ_COMPLETE, total (+ *add_in if given)];
 * d_tg_proto [K,256] = d(sum of the enabled losses)/d(tg_proto). */
int scan_transfer_proto(const float* tg_proto, const float* prototype, int32_t proto_iter, int32_t num_classes, int32_t flags,
                        const float* add_in, float* losses4, float* d_tg_proto, void* stream);

/* ---- K2: DBSCAN target-node sampling (loss.py:397-423 DBSCAN_batch_cpu; sklearn.cluster.DBSCAN
 *      eps=DBSCAN_EPS, min_samples=5, euclidean, brute) ------------------------------------------------
 * One level per call.  act_nchw [N, K, H, W] (channel 0 = background), rows_level = first row of the
 * level in `rows` [R,256].  Steps (all on device, no host sync):
 *   select   entries (n, cls>=1, y, x) with act > thr in flat order ((n*CLS+cls-1)*H+y)*W+x;
 *   points   p_i = fl32(rows[g_i, :] * act_i)                                  (n_points x 256, workspace);
 *   neighbours  d2 = |p_i|^2 + |p_j|^2 - 2 p_i.p_j evaluated in fp32 tiles, pairs within a band around
 *            eps^2 re-evaluated in fp64 exactly as sklearn does (float64 accumulation of fp32 inputs);
 *   core = #neighbours(incl. self) >= min_samples; union-find over core-core edges; cluster id = rank of
 *   the component's smallest index; border point -> smallest cluster id among its core neighbours; else -1.
 *   location mask: entry value = 1 for noise, 0 for cluster 0, id otherwise (loss.py:417-418); a location
 *   is positive iff any of its class entries is non-zero (loss.py:420-421); clustering is skipped (every
 *   selected entry positive) iff all selected points are exactly zero (loss.py:415).
 * Outputs: pos_mask [N*H*W] uint8, plabel [N*H*W] int64 = argmax_{c>=1} act + 1 (loss.py:500),
 *          point_labels int32 [cap] (sklearn labels, for tests), info (device int32[8]):
 *          [0] n_points, [1] n_clusters, [2] n_noise, [3] skipped, [4] error (1 = cap exceeded), [5] n_recheck.
 * workspace: scan_dbscan_workspace_bytes(cap). */
int64_t scan_dbscan_workspace_bytes(int64_t cap_points);
int scan_dbscan_level(const float* rows_level, const float* act_nchw, int32_t n_images, int32_t num_classes,
                      int32_t h, int32_t w, float thr, double eps, int32_t min_samples, int32_t cap_points,
                      uint8_t* pos_mask, int64_t* plabel, int32_t* point_labels, int32_t* info,
                      void* workspace, int64_t workspace_bytes, void* stream);
/* Stand-alone clustering of an explicit point set [n, dim] fp32 (dim % 4 == 0): labels int32 [n]. */
int scan_dbscan_points(const float* points, int32_t n, int32_t dim, double eps, int32_t min_samples,
                       int32_t* labels, int32_t* info, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* ---- K5: sigmoid focal loss (csrc/cuda/SigmoidFocalLoss_cuda.cu:21-101, the `_C.sigmoid_focalloss_*`
 *      pair wrapped by layers/sigmoid_focal_loss.py:9-37) and TEST.MODE ensembling (fcos.py:162-169) -- */
int scan_sigmoid_focal_fwd(const float* logits, const int32_t* targets, int64_t n_rows, int32_t num_classes,
                           float gamma, float alpha, float* losses, void* stream);
int scan_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses,
                           int64_t n_rows, int32_t num_classes, float gamma, float alpha, float* d_logits,
                           void* stream);
/* TEST.MODE ensembling of ALL levels in one launch.  mode 0 'common': out = sigmoid(cls); 1 'light': out = act[:,1:] (callers
 * that can hand out a view do not need the kernel); 2 'precision': 0.5*sigmoid(cls)+0.5*act[:,1:].
 * Per level l: cls [N, K-1, H_l, W_l] (array may be NULL for light), act [N, K, H_l, W_l], out [N, K-1, H_l, W_l];
 * the three pointer arrays are HOST arrays of device pointers, like scan_pack_rows'. */
int scan_ensemble_levels(const scan_levels_t* levels, const void* const* cls_logits_host, const void* const* act_host,
                         int32_t num_classes, int32_t mode, void* const* out_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCAN_B200_H_ */
