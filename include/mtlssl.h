/* mtlssl.h — C ABI of libmtlssl.so: the B200 (sm_100a) kernels of the multi-task detection
 * training path of wonheeML/mtl-ssl.
 *
 * The reference has NO native code and no FFI: every op below is a chain of TensorFlow-1.7 ops
 * driven from Python (graph mode).  Each entry point therefore cites the reference Python that
 * it replaces (paths relative to /root/reference/); INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host";
 *   - caller owns all buffers; kernels are stateless, stream-ordered, never synchronise;
 *   - return 0 on success, negative error code otherwise (mtl_last_error_string() explains);
 *   - activations NHWC bf16, weights [K,R,S,C] bf16, geometry / losses / logits fp32,
 *     indices int32; boxes are float4 [ymin, xmin, ymax, xmax];
 *   - `stream` is a cudaStream_t passed as void* so that this header needs no CUDA include.
 */
#ifndef MTLSSL_H_
#define MTLSSL_H_
#ifdef __cplusplus
extern "C" {
#endif

#define MTL_OK 0
#define MTL_ERR_ARG (-1)
#define MTL_ERR_CUDA (-2)
#define MTL_ERR_UNSUPPORTED (-3)

typedef void* mtl_stream_t;

const char* mtl_last_error_string(void);
int mtl_abi_version(void);
int mtl_device_sm_count(void);
long long mtl_launch_count(void);   /* kernels launched (or graph-captured) by this library so far */

/* ---- tcgen05 implicit-GEMM convolution engine (csrc/gemm_tc.cu) -------------------------
 * Replaces slim.conv2d / slim.fully_connected and their gradients
 * (slim/nets/resnet_v1.py:107-126, slim/nets/resnet_utils.py:77-122,
 *  object_detection/core/box_predictor.py:482-496, 728-746). */
typedef struct mtl_conv_args {
  int mode;                 /* 0 fprop, 1 dgrad, 2 wgrad */
  int N, H, W, C;           /* input-side activation x[N,H,W,C] */
  int K;                    /* output channels */
  int R, S, stride, pad_h, pad_w, dil;
  int P, Q;                 /* output-side spatial dims y[N,P,Q,K] */
  const void* x;            /* fprop/wgrad: x bf16 */
  const void* w;            /* fprop/dgrad: w[K,R,S,C] bf16 */
  const void* dy;           /* dgrad/wgrad: dy[N,P,Q,K] bf16 */
  void* out;                /* fprop: y; dgrad: dx; wgrad: dw fp32 [K,R,S,C] (accumulated) */
  int out_fp32;
  const float* bias;        /* fprop: per-K bias or NULL */
  const float* rowscale;    /* wgrad: per-K scale (frozen BN fold) or NULL */
  const void* res; int res_fp32;   /* added before relu/mask; same shape as out */
  const void* mask;         /* bf16, same shape as out: out = mask > 0 ? out : 0 */
  int relu;                 /* 0 none, 1 ReLU, 2 ReLU6 */
  float alpha;              /* wgrad scale */
  int force_bn;             /* 0 = auto tile width */
  int force_splits;         /* 0 = auto split-K */
  float mask_hi;            /* dgrad: > 0 -> ReLU6 mask: gradient only where 0 < mask < mask_hi */
  long long dy_ld, out_ld, res_ld, mask_ld;   /* pixel pitches in elements for channel slices; 0 = dense */
  float bias_scale;         /* bias added as bias[n] * bias_scale; 0 = 1.0 */
  int force_stages;         /* 0 = auto operand pipeline depth */
  void* ws;                 /* fprop/dgrad split-K workspace (zero filled; left zeroed) or NULL = never split */
  long long ws_bytes;
  int force_cluster;        /* 0 auto, 1 never, 2 always: CTA pairs sharing multicast weight tiles (fprop/dgrad) */
  int max_ctas;             /* 0 = whole GPU; else size the persistent grid for this many SMs (overlapped side work) */
  float* pool_out;          /* fprop with residual + ReLU, forward-only tails (core/box_predictor.py:470-476 reduce_mean over
                             * the ROI grid follows): y is NOT stored; instead pool_out[(2g + s) * K + k] = sum of y[row, k]
                             * (bf16-rounded) over the rows of 32-row group g that lie in its first (s = 0) / second (s = 1)
                             * window of pool_hw consecutive rows; [2 * ceil(N*P*Q / 32), K] fp32; NULL = store y */
  int pool_hw;              /* rows per pooling window (P*Q of one ROI), >= 32 */
} mtl_conv_args;
int mtl_conv_tc(const mtl_conv_args* args /* host */, mtl_stream_t stream);
/* bytes of zeroed workspace mtl_conv_tc would use to split the K loop of this fprop/dgrad (0 = no split) */
long long mtl_conv_tc_ws_bytes(const mtl_conv_args* args /* host */);
/* Grouped launch: n INDEPENDENT problems of one kernel instance (equal mtl_conv_tc_group_key) in one persistent grid
 * whose CTAs walk the concatenated tile space (the trunk's 81 weight-gradient GEMMs at batch 1 become a few launches).
 * build() plans the problems into a table in HOST memory (n x entry_bytes, 128-byte aligned; info[4] = total tiles, tile
 * width, mode, operand path; K of a weight-gradient problem is cut into pieces of ~target_k_iters 64-deep steps, 0 =
 * never); the caller copies the table to device memory once and keeps both alive; launch() only enqueues. */
long long mtl_conv_tc_group_entry_bytes(void);
long long mtl_conv_tc_group_key(const mtl_conv_args* args /* host */);
int mtl_conv_tc_group_build(const mtl_conv_args* args /* host */, int n, int target_k_iters, void* host_table /* host */,
                            int* info /* host */);
int mtl_conv_tc_group_launch(const void* host_table /* host */, const void* dev_table, int n, const int* info /* host */,
                             int max_ctas, mtl_stream_t stream);

/* ---- anchors (object_detection/anchor_generators/grid_anchor_generator.py:96-214) ------ */
int mtl_grid_anchors(int Hf, int Wf, const float* scales /* host */, int num_scales,
                     const float* aspect_ratios /* host */, int num_aspect_ratios, float base_h, float base_w,
                     float stride_h, float stride_w, float off_h, float off_w, float* anchors /* [Hf*Wf*A,4] */,
                     mtl_stream_t stream);
/* core/box_list_ops.py:140-169 prune_outside_window (ordered). */
int mtl_prune_outside_window(const float* boxes, int N, float wy0, float wx0, float wy1, float wx1,
                             int* keep_idx /* [N] */, float* kept_boxes /* [N,4] */, int* num_keep /* [1] */,
                             mtl_stream_t stream);

/* ---- RPN proposals (meta_architectures/faster_rcnn_meta_arch.py:1055-1115,
 *      box_coders/faster_rcnn_box_coder.py:92-118, core/post_processing.py:25-164) ------- */
int mtl_rpn_decode(const float* rpn_out /* [B,HW,ld] */, long long ld, int box_col0, int cls_col0, int A, int HW,
                   const int* keep_idx /* [Nk] or NULL */, const float* anchors /* [Nk,4] */, int Nk, int B,
                   float img_h, float img_w, float score_thresh, float* boxes /* [B,Nk,4] clipped */,
                   float* scores /* [B,Nk] */, unsigned long long* keys /* [B,Nk] */, mtl_stream_t stream);
int mtl_nms_make_keys(const float* boxes, const float* scores, int B, int N, float score_thresh, int require_area,
                      unsigned long long* keys, mtl_stream_t stream);
int mtl_rank_sort_desc(const unsigned long long* keys, int B, int N, int* order /* [B,N] */,
                       int* num_valid /* [B] */, int* rank_ws /* [B,N] workspace */, mtl_stream_t stream);
/* tf.image.non_max_suppression semantics (TF 1.7), one block per image, zero padded output. */
int mtl_nms(const float* boxes /* [B,N,4] */, const float* scores /* [B,N] or NULL */, const int* order,
            const int* num_valid, int B, int N, float iou_thresh, int max_out, float* out_boxes /* [B,max,4] */,
            float* out_scores /* [B,max] or NULL */, int* out_idx /* [B,max] or NULL */, int* num_out /* [B] */,
            mtl_stream_t stream);

/* ---- inference-time proposal path (fmA:586-590 anchors clipped, not pruned; fmA:1111-1131 the NMS output is
 *      the proposal set, no minibatch sampling) ------------------------------------------------------------ */
int mtl_clip_boxes(const float* boxes /* [N,4] */, int N, float wy0, float wx0, float wy1, float wx1,
                   float* out /* [N,4] */, mtl_stream_t stream);
int mtl_proposals_from_nms(const float* nms_boxes /* [B,M,4] */, const float* nms_scores /* [B,M] */,
                           const int* nms_num /* [B] */, int B, int M, float img_h, float img_w,
                           float* out_abs /* [B,M,4] */, float* out_norm /* [B,M,4] */,
                           float* out_scores /* [B,M] or NULL */, int* num_out /* [B] */, mtl_stream_t stream);

/* ---- second-stage detections (meta_architectures/faster_rcnn_meta_arch.py:1387-1469,
 *      core/post_processing.py:25-312).  mtl_detection_decode produces the per-(image, class) NMS inputs in
 *      class-major layout [B,K,P]; mtl_rank_sort_desc + mtl_nms run with B*K as their batch; the two functions
 *      below merge the per-class survivors and keep the top max_total detections (zero padded). ------------- */
int mtl_detection_decode(const float* box_encodings /* [B*P,K,4] */, const float* class_logits /* [B*P,K+1] */,
                         const float* proposals /* [B,P,4] abs */, const int* num_proposals /* [B] */, int B, int P,
                         int K, float img_h, float img_w, float score_thresh, int score_mode /* 0 id, 1 softmax, 2 sigmoid */,
                         float* boxes_norm /* [B,K,P,4] clipped, window frame */, float* scores /* [B,K,P] */,
                         unsigned long long* keys /* [B,K,P] */, float* decoded_abs /* [B,K,P,4] or NULL */,
                         mtl_stream_t stream);
int mtl_detection_merge_keys(const float* cls_scores /* [B,K,M] */, const int* cls_num /* [B,K] */, int B, int K, int M,
                             unsigned long long* keys /* [B,K*M] */, mtl_stream_t stream);
int mtl_detection_gather(const float* cls_boxes /* [B,K,M,4] */, const float* cls_scores /* [B,K,M] */,
                         const int* order /* [B,K*M] */, const int* num_valid /* [B] */, int B, int K, int M, int T,
                         float* det_boxes /* [B,T,4] */, float* det_scores /* [B,T] */, float* det_classes /* [B,T] */,
                         float* num_detections /* [B] */, mtl_stream_t stream);

/* ---- matching / sampling / targets (core/region_similarity_calculator.py:57-74,
 *      matchers/argmax_matcher.py:102-175, core/target_assigner.py:99-213,
 *      core/balanced_positive_negative_sampler.py:50-91) ----------------------------------- */
int mtl_iou_match(const float* gt /* [B,Gmax,4] */, const int* num_gt /* [B] */, int Gmax,
                  const float* boxes /* [B or 1, N, 4] */, long long box_batch_stride /* in boxes, 0 = shared */,
                  const int* num_boxes /* [B] or NULL */, int B, int N, float matched_thr, float unmatched_thr,
                  int use_thresholds, int force_match, int* match /* [B,N]: >=0 row, -1 neg, -2 ignore, -3 absent */,
                  float* max_iou /* [B,N] or NULL */, unsigned long long* row_best /* [B,Gmax] workspace */,
                  mtl_stream_t stream);
int mtl_balanced_sample(const int* match, const float* keys /* [B,N] explicit shuffle keys */, int B, int N,
                        int batch_size, float positive_fraction, unsigned char* sampled /* [B,N] */,
                        int* counts /* [B,4]: #pos cand, #neg cand, #pos taken, #taken */, mtl_stream_t stream);
int mtl_gather_sampled(const float* boxes, const float* scores, const unsigned char* sampled, int B, int N, int P,
                       float img_h, float img_w, float* out_abs /* [B,P,4] */, float* out_norm /* [B,P,4] */,
                       float* out_scores /* [B,P] or NULL */, int* num_out /* [B] */, mtl_stream_t stream);
int mtl_detection_targets(const int* match, const float* proposals /* [B,P,4] abs */, const float* gt,
                          const int* gt_classes /* [B,Gmax] 1..K */, const float* gt_closeness /* [B,Gmax,K1] or NULL */,
                          int B, int Gmax, int P, int K1, int* cls_targets, float* reg_targets, float* reg_weights,
                          float* cls_weights, float* closeness_targets, float* closeness_weights,
                          mtl_stream_t stream);
int mtl_rpn_targets(const int* match, const float* anchors, const float* gt, int B, int Gmax, int N,
                    float* cls_targets, float* cls_weights, float* reg_targets, float* reg_weights,
                    mtl_stream_t stream);
/* meta_architectures/faster_rcnn_meta_arch.py:783-803 (refine head window expansion). */
int mtl_expand_windows(const float* proposals_norm /* [B,P,4] */, int B, int P, int n_expand,
                       float* out /* [n_expand+1,B,P,4] */, int* box_ind /* same leading dims or NULL */,
                       mtl_stream_t stream);

/* ---- ROI / pooling / stem (tf.image.crop_and_resize @ fmA:1340; utils/ops.py:462-609;
 *      slim max_pool2d; box_predictor.py:470-472; fe preprocess :74-90) ---------------------- */
int mtl_crop_and_resize_fwd(const void* feat /* bf16 [B,H,W,C] */, int B, int H, int W, int C,
                            const float* boxes /* [R,4] normalised */, const int* box_ind /* [R] or NULL = 0 */,
                            int R, int crop_h, int crop_w, void* out /* bf16 [R,ch,cw,C] */, mtl_stream_t stream);
int mtl_crop_and_resize_bwd(const void* dcrop /* bf16 */, int B, int H, int W, int C, const float* boxes,
                            const int* box_ind, int R, int crop_h, int crop_w, float* dfeat /* fp32, += */,
                            mtl_stream_t stream);
int mtl_maxpool_fwd(const void* x, int N, int H, int W, int C, int k, int stride, int pad_h, int pad_w, int P,
                    int Q, void* y, long long ldy /* pixel pitch of y, 0 = C */, mtl_stream_t stream);
int mtl_maxpool_bwd(const void* x, const void* dy, long long ldy, int N, int H, int W, int C, int k, int stride,
                    int pad_h, int pad_w, int P, int Q, void* dx, mtl_stream_t stream);
/* slim.avg_pool2d(3, stride 1, SAME) of Mixed_5b (slim/nets/inception_resnet_v2.py:176-181); backward=1: gradient */
int mtl_avgpool3x3_same(const void* x, int N, int H, int W, int C, int backward, void* y, mtl_stream_t stream);
int mtl_avgpool_fwd(const void* x /* bf16 [R,HW,C] */, int R, int HW, int C, void* y /* bf16 [R,C] */,
                    mtl_stream_t stream);
int mtl_avgpool_bwd(const void* dy, int dy_fp32, long long ldy, const void* relu_mask /* bf16 [R,HW,C] or NULL */,
                    float mask_hi /* > 0: ReLU6-style upper bound */, int R, int HW, int C,
                    void* dx /* bf16 [R,HW,C] */, mtl_stream_t stream);
/* Fused MaskRCNNBoxPredictor head (core/box_predictor.py:470-500, :568-602): spatial average over the ROI grid + the
 * fully connected layer(s) in one kernel per ROI batch; w = the head's weight rows [n, C] bf16 (box and class matrices
 * back to back), out = fp32 logits [R, ldo]. */
int mtl_head_fwd(const void* x /* bf16 [R,HW,C] */, int R, int HW, int C, const void* w /* bf16 [n,C] */,
                 const float* bias /* [n] or NULL */, int n, void* pooled /* bf16 [R,C] out */,
                 float* out /* [R,ldo] */, long long ldo, mtl_stream_t stream);
/* mtl_head_fwd on the partial row sums mtl_conv_tc wrote with pool_out (the tail's last conv never stores its output) */
int mtl_head_fwd_pooled(const float* part /* [2*ceil(R*HW/32), C] */, int R, int HW, int C, const void* w /* [n,C] bf16 */,
                        const float* bias /* [n] or NULL */, int n, void* pooled /* [R,C] bf16 */,
                        float* out /* [R,ldo] */, long long ldo, mtl_stream_t stream);
/* Its backward: bf16 copy of the logit gradient (operand of the weight-gradient GEMM), bias gradient (+=), and -- dx not
 * NULL -- the gradient w.r.t. x: (d_out x w) / HW broadcast over the grid, zero where x <= 0 (or >= mask_hi > 0). */
int mtl_head_bwd(const float* d_out /* [R,ldd] */, long long ldd, int n, const void* w /* bf16 [n,C] */,
                 const void* x /* bf16 [R,HW,C] ReLU mask or NULL */, float mask_hi, int R, int HW, int C,
                 void* dyb /* bf16 [R,n] out */, float* db /* [n] += or NULL */, void* dx /* bf16 [R,HW,C] or NULL */,
                 mtl_stream_t stream);

int mtl_im2col_f32(const float* img /* [B,H,W,C<=4] */, int B, int H, int W, int C, int R, int S, int stride,
                   int pad_h, int pad_w, int P, int Q, const float* mean /* host [C] or NULL */, float scale,
                   void* out /* bf16 [B*P*Q, ld] */, int ld, mtl_stream_t stream);
/* core/preprocessor.py:1362-1419 resize_to_range -> tf.image.resize_images (bilinear). */
int mtl_resize_bilinear_f32(const float* x, int B, int H, int W, int C, int out_h, int out_w, float* y,
                            mtl_stream_t stream);
int mtl_preprocess(const float* img, long long total, int C, const float* mean /* host */, float scale, float* out,
                   mtl_stream_t stream);
/* slim.separable_conv2d depthwise stage (slim/nets/mobilenet_v1.py:230-238); w bf16 [C,3,3];
 * activation 0 none / 1 ReLU / 2 ReLU6; dgrad masks by (0 < mask [< mask_hi]). */
int mtl_dwconv3x3_fwd(const void* x, const void* w, const float* bias, int N, int H, int W, int C, int stride,
                      int pad_h, int pad_w, int P, int Q, int activation, void* y, mtl_stream_t stream);
int mtl_dwconv3x3_dgrad(const void* dy, const void* w, int N, int H, int W, int C, int stride, int pad_h, int pad_w,
                        int P, int Q, const void* mask, float mask_hi, void* dx, mtl_stream_t stream);
int mtl_dwconv3x3_wgrad(const void* dy, const void* x, int N, int H, int W, int C, int stride, int pad_h, int pad_w,
                        int P, int Q, const float* scale, float* dw /* fp32 [C,3,3], += */, mtl_stream_t stream);
int mtl_psroi_fwd(const void* feat /* bf16 [B,H,W,Ct] */, int B, int H, int W, int Ct, int c0 /* first PS channel */,
                  int D /* outputs per bin */, int nby, int nbx, int crop_h, int crop_w, const float* boxes,
                  const int* box_ind, int R, float* out /* [R,ldo] at column ocol0 */, long long ldo, int ocol0,
                  mtl_stream_t stream);
int mtl_psroi_bwd(const float* dout, long long ldo, int ocol0, int B, int H, int W, int Ct, int c0, int D, int nby,
                  int nbx, int crop_h, int crop_w, const float* boxes, const int* box_ind, int R,
                  float* dfeat /* fp32 [B,H,W,Ct], += */, mtl_stream_t stream);
int mtl_add_bf16_to_f32(const void* src /* bf16 */, long long n, float* dst /* += */, mtl_stream_t stream);

/* ---- fused loss epilogues (core/losses.py:169-196, 285-352; fmA:1591-1881) ---------------- */
int mtl_rpn_loss(const float* rpn_out, long long ld, int box_col0, int cls_col0, int A, int HW, const int* keep_idx,
                 const float* anchors, int Nk, const float* gt, int Gmax, const int* match,
                 const unsigned char* sampled, const int* counts, int B, float loc_weight, float obj_weight,
                 float sigma, float* losses /* [2], += */, void* d_rpn_out /* bf16 [B,HW,ld] or NULL */,
                 mtl_stream_t stream);
int mtl_box_classifier_loss(const float* head_out, long long ld, int box_col0, int cls_col0, int K,
                            const int* cls_targets, const float* reg_targets, const float* reg_weights,
                            const float* cls_weights, const int* num_proposals, int B, int P, float loc_weight,
                            float cls_weight, float* losses /* [2], += */, float* d_head /* or NULL */,
                            long long ldd, mtl_stream_t stream);
int mtl_softmax_ce(const float* logits, long long ld, int col0, int C, const float* soft_targets, long long ldt,
                   int tcol0, const int* hard_targets, const float* weights, const int* num_proposals, int P,
                   int per_image_norm, long long rows, float scale, float* loss /* [1], += */, float* d_logits,
                   long long ldd, int dcol0, int accumulate, mtl_stream_t stream);
/* core/mask_predictor.py:90-119 + fmA:1860-1881 */
int mtl_edgemask_fwd(const void* feat /* bf16 [npix,C] */, long long npix, int C, const float* w /* [2,C] */,
                     const float* bias /* [2] */, float* act /* [npix,2] tanh */, mtl_stream_t stream);
int mtl_edgemask_loss(const float* act, int B, int H, int W, const float* gt /* [B,2,oh,ow] */, int oh, int ow,
                      float loss_weight, float* loss /* += */, float* d_act /* [B,H,W,2] or NULL */,
                      mtl_stream_t stream);
int mtl_edgemask_bwd(const void* feat, long long npix, int C, const float* w, const float* act, const float* d_act,
                     float* dfeat /* fp32 += or NULL */, float* dw /* [2,C] += */, float* dbias /* [2] += */,
                     mtl_stream_t stream);
/* fmA:764-846 refiner */
int mtl_fc_fwd(const float* x, long long ldx, const float* w /* [N,K] */, const float* bias, const float* res,
               long long ldr, int M, int N, int K, float* y, long long ldy, mtl_stream_t stream);
int mtl_fc_bwd(const float* x, long long ldx, const float* w, const float* dy, long long ldy, int M, int N, int K,
               float* dw /* += */, float* db /* += or NULL */, float* dx /* or NULL */, long long lddx,
               mtl_stream_t stream);
int mtl_refine_concat(const float* org, long long ldo, int ocol0, const float* win, long long ldw, int wcol0, int E,
                      const float* close, long long ldc, int ccol0, int rows, int K1, float* out, long long ldout,
                      mtl_stream_t stream);
int mtl_colsum(const void* dy, int is_fp32, long long ld, long long M, int N, float alpha, float* db /* += */,
               mtl_stream_t stream);

/* ---- optimizer (slim/learning.py:282-301; builders/optimizer_builder.py:49-53;
 *      slim/deployment/model_deploy.py:198-307) --------------------------------------------- */
typedef struct mtl_tensor_desc {
  long long offset, numel, row_len, scale_off;
  float l2_weight, grad_mult;
  int trainable, pad_;
} mtl_tensor_desc;
typedef struct mtl_chunk_desc {
  int tensor, len;
  long long start;
} mtl_chunk_desc;
int mtl_opt_chunk_size(void);
int mtl_opt_stats(const mtl_tensor_desc* tensors, int num_tensors, const mtl_chunk_desc* chunks, int num_chunks,
                  const float* params, const float* grads, float grad_scale, float* stats /* [T,2] */,
                  float* reg_loss /* [1] or NULL */, const int* chunk_start /* [T+1] first chunk of every tensor */,
                  float* partials /* [num_chunks,2]: fixed-order (run-to-run and replica-to-replica identical) sums,
                                     NULL = fp32 atomics */,
                  mtl_stream_t stream);
/* the same statistics for the tensors [t0, t1) only (`chunks` = first chunk of tensor t0): lets the update of a
 * finished gradient bucket overlap the rest of the backward pass; mtl_opt_reg_loss then sums the L2 terms. */
int mtl_opt_stats_range(const mtl_tensor_desc* tensors, int t0, int t1, const mtl_chunk_desc* chunks, int num_chunks,
                        const float* params, const float* grads, float grad_scale, float* stats /* [T,2] */,
                        const int* chunk_start /* [T+1] */, int chunk0 /* index of chunks[0] */,
                        float* partials /* [all chunks,2] or NULL */, mtl_stream_t stream);
int mtl_opt_reg_loss(const mtl_tensor_desc* tensors, int num_tensors, const float* stats, float* reg_loss /* [1] */,
                     mtl_stream_t stream);
int mtl_opt_apply(const mtl_tensor_desc* tensors, const mtl_chunk_desc* chunks, int num_chunks, float* params,
                  float* grads, float* momentum, void* params_bf16, const float* fold_scales, const float* stats,
                  const float* hyper /* [lr, momentum, clip_norm] */, float grad_scale, mtl_stream_t stream);
/* mtl_opt_apply over the tensors [t0, t1) (`chunks` = first chunk of tensor t0) that also refreshes stats[2t], the squared
 * norm of the UPDATED weights (fixed summation order through `partials`), so that the next step's regularisation loss needs
 * no extra pass over them; stats[2t + 1] is left at 0 until the next statistics pass. */
int mtl_opt_apply_norms(const mtl_tensor_desc* tensors, int t0, int t1, const mtl_chunk_desc* chunks, int num_chunks,
                        float* params, float* grads, float* momentum, void* params_bf16, const float* fold_scales,
                        float* stats /* [T,2] */, const float* hyper, float grad_scale, const int* chunk_start /* [T+1] */,
                        int chunk0 /* index of chunks[0] */, float* partials /* [all chunks,2] */, mtl_stream_t stream);
int mtl_opt_fold(const mtl_tensor_desc* tensors, const mtl_chunk_desc* chunks, int num_chunks, const float* params,
                 void* params_bf16, const float* fold_scales, mtl_stream_t stream);
int mtl_cast_f32_bf16(const float* a, long long n, float alpha, void* out, mtl_stream_t stream);
int mtl_relu_bwd_merge(const float* a, const void* b, const void* mask, long long n, void* out, mtl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MTLSSL_H_ */
