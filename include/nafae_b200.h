/*
 * nafae_b200.h -- C ABI of libnafae_b200.so: the B200-native (sm_100a) drop-in for the native
 * layer of jshi31/NAFAE's per-segment grounding hot path.
 *
 * Conventions (shared by every entry point)
 *   - plain C: raw DEVICE pointers, ints, floats and a cudaStream_t; no torch / THC types.
 *   - return value: 1 = ok (the reference's convention, e.g. roi_align_cuda.c:41),
 *                   0 = invalid argument (reference: `size_rois != 5` -> 0, roi_align_cuda.c:19-22),
 *                  <0 = -(cudaError_t) of a failed launch.  Never exit()s (the reference's
 *                   launchers call exit(-1), roi_align_kernel.cu:84-88), never prints.
 *     nafae_last_error() returns a thread-local, human readable message for the last 0 / <0.
 *   - stream ordered, asynchronous, no host synchronisation, no allocation: scratch memory is a
 *     caller-provided workspace whose size the matching *_workspace_bytes() function returns.
 *     Every entry point may be captured into a CUDA graph.
 *   - all tensors are dense, row-major ("contiguous" in torch terms), fp32 unless stated.
 *
 * Each declaration cites the reference interface (path:line under the reference tree) it
 * replaces.  INTEGRATION.md shows the binding a maintainer of the reference would add.
 */
#ifndef NAFAE_B200_H_
#define NAFAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t; /* same opaque handle as CUDA driver_types.h */
#endif

#define NAFAE_B200_ABI_VERSION 3 /* 3: nafae_allreduce_avg(flags), nafae_gate_sync, nafae_mc_*, nafae_allreduce_mc */

/* pooling applied on top of the sampled RoIAlign grid (modules/roi_align.py:6-42) */
#define NAFAE_POOL_NONE 0 /* RoIAlign    : output is the aligned_height x aligned_width grid   */
#define NAFAE_POOL_AVG 1  /* RoIAlignAvg : sample (h+1)x(w+1), avg_pool2d(kernel 2, stride 1)  */
#define NAFAE_POOL_MAX 2  /* RoIAlignMax : sample (h+1)x(w+1), max_pool2d(kernel 2, stride 1)  */

/* flags */
#define NAFAE_FLAG_NO_GATE 2u /* nafae_roi_align_forward: ignore the workspace's residency gate --
                                 nobody waits on this launch (keeps gate epochs paired with waiters) */
#define NAFAE_FLAG_OUT_BF16 4u /* nafae_roi_align_forward: top_data is bf16 -- the (R, C*7*7) row-major
                                  A operand of the bridge GEMM (nafae_gemm_bf16_tn) written directly,
                                  half the output bytes; bandwidth-kernel shapes only */
#define NAFAE_FLAG_OVERWRITE 8u /* nafae_roi_align_backward: bottom_diff is OVERWRITTEN (it need not be
                                   zero-filled): every cell is written exactly once */
#define NAFAE_FLAG_DETERMINISTIC 16u /* nafae_roi_align_backward (RoIAlignAvg 7x7): fixed summation
                                        order, bitwise reproducible (cell-gather kernel, no atomics);
                                        slower than the default shared-memory scatter */
#define NAFAE_FLAG_EXACT 1u /* reference-order arithmetic (mixed fp32/fp64 exactly as the
                               reference kernel evaluates it): bit-identical pooled features,
                               slower.  Default (0) = fp32 FMA path, <= 1e-4 relative. */

int nafae_abi_version(void);
const char* nafae_last_error(void);

/* Persistent kernels (the RoIAlign slab kernel) launch one CTA per SM.  When a collective runs
 * concurrently on another stream (data-parallel gradient all-reduce), leave `n` SMs free for its
 * CTAs so neither kernel waits for the other's residency.  Per CUDA device (the calling thread's
 * current one), default 0; returns the previous value.  n is clamped to [0, SMs-1]. */
int nafae_set_reserved_sms(int n);

/* Residency gate for kernels that run CONCURRENTLY with the persistent RoIAlign kernel (the head
 * of an earlier batch, the gradient all-reduce).  The block scheduler places CTAs breadth-first:
 * a concurrent kernel launched at the same time lands on every SM and keeps the 210 KB-per-CTA
 * persistent kernel off those SMs until its CTAs retire.  With a gate the persistent kernel bumps
 * an epoch once ALL its CTAs are resident, and nafae_gate_wait enqueues a one-warp kernel (it
 * co-resides with anything) that returns only then -- so whatever follows it on `stream` can only
 * land on the SMs left free by nafae_set_reserved_sms.  `gate` = the workspace passed to
 * nafae_roi_align_forward (its first NAFAE_GATE_BYTES); `slot` in [0, 6) identifies the waiting
 * branch (one slot per concurrent stream).  The wait returns once the gate has been opened by a
 * launch this slot has not yet seen. */
#define NAFAE_GATE_BYTES 32
#define NAFAE_ROI_ALIGN_WS_BYTES 64
int nafae_gate_wait(void* gate, int slot, cudaStream_t stream);
/* Marks every gated launch enqueued on `stream` so far as seen by every slot: the next
 * nafae_gate_wait of any slot blocks until the NEXT gated launch opens the gate.  Enqueue it once
 * after gated launches that had no waiter (warm-up runs), before the paired launches begin --
 * otherwise each wait would be satisfied by the previous launch's open. */
int nafae_gate_sync(void* gate, cudaStream_t stream);

/* Diagnostics (tests): `num_ctas` CTAs that each hold `smem_bytes` of shared memory -- a whole SM when
 * close to 227 KB -- and do nothing for `nanoseconds` (<= 2 s) of wall time.  Lets a test squeeze the
 * path's kernels onto the few SMs that remain (forward-progress checks). */
int nafae_debug_occupy_sms(int num_ctas, int smem_bytes, unsigned long long nanoseconds,
                           cudaStream_t stream);

/* ------------------------------------------------------------------------------- NMS ---- */

/* Replaces nms_cuda_compute() -- lib/model/nms/src/nms_cuda_kernel.h:5-6 (impl
 * nms_cuda_kernel.cu:87-161) and its THC glue nms_cuda() -- lib/model/nms/src/nms_cuda.h:4-5.
 * Same symbol, same argument meaning: boxes (boxes_num, boxes_dim>=4) rows [x1,y1,x2,y2,...]
 * sorted by score descending, DEVICE memory (the reference's "boxes_host" is a device pointer
 * too, nms_cuda.c:12-14); keep_out (boxes_num) int32 and num_out (1) int32 DEVICE buffers owned
 * by the caller.  Greedy, IoU with the +1 pixel convention, strict '>' (nms_cuda_kernel.cu:31-39,
 * 78).  Differences: no cudaMalloc/cudaFree, no D2H mask copy, no host sweep -- runs on the
 * legacy default stream (like the reference's <<<blocks, threads>>>) using an internal
 * per-device scratch cache, and returns without synchronising. */
void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num, int boxes_dim,
                      float nms_overlap_thresh);

/* Batched, stream-ordered form of the same operator: F independent frames in one call.
 * boxes (F, n, boxes_dim); keep_out (F, n) int32; num_out (F) int32.
 * workspace: nafae_nms_workspace_bytes(F, n) bytes of device memory. */
size_t nafae_nms_workspace_bytes(int num_frames, int boxes_num);
int nafae_nms_batched(int* keep_out, int* num_out, const float* boxes, int num_frames,
                      int boxes_num, int boxes_dim, float nms_overlap_thresh, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream);

/* Replaces the part of _ProposalLayer.forward BEFORE its per-frame loop --
 * lib/model/rpn/proposal_layer.py:66-125: anchor enumeration (base windows of generate_anchors.py +
 * feature-stride shifts, :80-93), the (H, W, A) re-ordering of the NCHW RPN outputs (:98-103),
 * bbox_transform_inv (lib/model/rpn/bbox_transform.py:77-103), clip_boxes (:125-133) and the per-frame
 * torch.sort(scores, 1, True) (:125) -- in two launches for the whole batch.
 *   rpn_cls_prob (B, 2A, H, W): channels [A, 2A) are the foreground probabilities (:66)
 *   rpn_bbox_pred (B, 4A, H, W); im_info (B, 3) = [height, width, scale] DEVICE; anchors (A, 4) DEVICE
 *   = generate_anchors(scales, ratios) as float32
 * Outputs, ready for nafae_proposal_tail: proposals_sorted (B, m, 4) and scores_sorted (B, m) in
 * score-descending order per frame, m = pre_nms_topn when 0 < pre_nms_topn < B*H*W*A (the reference
 * compares with the element count of the whole batch, :139) and < H*W*A, else H*W*A; order (B, m) int32
 * = the anchor index (h*W*A + w*A + a) of every sorted position, or NULL.  Equal scores keep ascending
 * anchor index (a stable sort; the reference leaves their order unspecified).  H*W*A <= 25600 (the sort
 * lives in shared memory).  workspace: nafae_proposal_front_workspace_bytes(...) bytes. */
size_t nafae_proposal_front_workspace_bytes(int batch_size, int num_anchors, int height, int width);
int nafae_proposal_front(const float* rpn_cls_prob, const float* rpn_bbox_pred, const float* im_info,
                         const float* anchors, int batch_size, int num_anchors, int height, int width,
                         float feat_stride, int pre_nms_topn, float* proposals_sorted, float* scores_sorted,
                         int* order, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Replaces the per-frame Python loop of _ProposalLayer.forward --
 * lib/model/rpn/proposal_layer.py:127-163 (slice to pre_nms_topN :139-140, nms() :150, first
 * post_nms_topN keeps :154-155, zero padding + frame index in column 0 :158-163) -- for all
 * frames in ONE launch, stopping each frame's greedy scan as soon as post_nms_topn boxes are
 * kept.  proposals (F, n, 4) and scores (F, n) must already be in score-descending order per
 * frame (the torch.sort of :125).  Outputs are fully written (padding included):
 *   rois (F, post_nms_topn, 5) = [frame, x1, y1, x2, y2], roi_scores (F, post_nms_topn),
 *   num_kept (F) int32 or NULL.  post_nms_topn <= 0 is invalid here (use nafae_nms_batched). */
int nafae_proposal_tail(const float* proposals, const float* scores, int num_frames, int boxes_num,
                        int pre_nms_topn, int post_nms_topn, float nms_thresh, float* rois,
                        float* roi_scores, int* num_kept, cudaStream_t stream);

/* -------------------------------------------------------------------------- RoIAlign ---- */

/* Replace ROIAlignForwardLaucher / ROIAlignBackwardLaucher --
 * lib/model/roi_align/src/roi_align_kernel.h:13-27 (impl roi_align_kernel.cu:73-91, 145-162);
 * same symbols, same signatures.  bottom_data (B, C, H, W) NCHW; bottom_rois (R, 5) =
 * [batch_index, x1, y1, x2, y2] in image coordinates; top_data (R, C, ah, aw).  Backward
 * ACCUMULATES into bottom_diff, which the caller zero-fills (functions/roi_align.py:38-39).
 * These two always use the reference-order arithmetic (NAFAE_FLAG_EXACT). */
int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                           const int height, const int width, const int channels,
                           const int aligned_height, const int aligned_width,
                           const float* bottom_rois, float* top_data, cudaStream_t stream);
int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                            const int num_rois, const int height, const int width,
                            const int channels, const int aligned_height, const int aligned_width,
                            const float* bottom_rois, float* bottom_diff, cudaStream_t stream);

/* Fused module-level operator: RoIAlign / RoIAlignAvg / RoIAlignMax.forward --
 * lib/model/roi_align/modules/roi_align.py:14-16, 26-29, 39-42 -- i.e. roi_align_forward_cuda
 * (src/roi_align_cuda.h:1-2) plus the following avg_pool2d / max_pool2d, without the
 * (R, C, h+1, w+1) intermediate.  out_height x out_width is the MODULE's aligned size (7x7 for
 * RoIAlignAvg(7, 7, 1/16)); with pool_mode != NONE the sampled grid is (out+1) x (out+1).
 * top_data (R, C, out_height, out_width) is fully written (no zero-fill needed).
 * workspace: optional (NULL / 0 is fine).  nafae_roi_align_workspace_bytes(batch_size, num_rois)
 * (= NAFAE_ROI_ALIGN_WS_BYTES) bytes of device memory, ZERO-INITIALISED ONCE by the caller and then
 * left to the library, private to one stream: the residency gate of nafae_gate_wait. */
size_t nafae_roi_align_workspace_bytes(int batch_size, int num_rois);
/* CTAs the persistent RoIAlign kernel launches for `num_units` (frame, 8-channel-group) work units
 * under the current nafae_set_reserved_sms setting: the smallest grid whose busiest CTA has no more
 * units than with every available SM (reporting / capacity planning; no launch). */
int nafae_roi_align_persistent_ctas(int num_units);
int nafae_roi_align_forward(const float* bottom_data, float spatial_scale, int batch_size,
                            int num_rois, int height, int width, int channels, int out_height,
                            int out_width, int pool_mode, const float* bottom_rois, void* top_data,
                            unsigned flags, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream);

/* Backward of the fused operator (roi_align_backward_cuda, src/roi_align_cuda.h:4-5, preceded by
 * the pool's backward that autograd derives).  top_diff (R, C, out_height, out_width);
 * bottom_data is only read for NAFAE_POOL_MAX (may be NULL otherwise).  ACCUMULATES into
 * bottom_diff (B, C, H, W), which the caller zero-fills (functions/roi_align.py:38-39) -- unless
 * NAFAE_FLAG_OVERWRITE is set: then bottom_diff is fully written by the call (no zero-fill needed).
 * RoIAlignAvg 7x7 (the configuration the reference instantiates) never sends an atomic to HBM: a CTA
 * owns a (frame, 8-channel) slab of bottom_diff in shared memory, scatters the frame's RoIs into it and
 * stores (OVERWRITE) or reduce-adds (default) the slab with bulk async copies.  Like the reference's
 * atomicAdd the summation order is not fixed; NAFAE_FLAG_DETERMINISTIC selects a cell-gather kernel
 * with a fixed order instead.  NAFAE_FLAG_EXACT and every other shape use the reference-style
 * global-atomic kernel. */
int nafae_roi_align_backward(const float* top_diff, const float* bottom_data, float spatial_scale,
                             int batch_size, int num_rois, int height, int width, int channels,
                             int out_height, int out_width, int pool_mode, const float* bottom_rois,
                             float* bottom_diff, unsigned flags, cudaStream_t stream);

/* --------------------------------------------------------------------------- RoIPool ---- */

/* Replace ROIPoolForwardLaucher / ROIPoolBackwardLaucher --
 * lib/model/roi_pooling/src/roi_pooling_kernel.h:8-18 (impl roi_pooling_kernel.cu:95-125,
 * 205-234); same symbols, same signatures.  argmax_data (R, C, ph, pw) int32 holds the flat index
 * into the whole (B, C, H, W) batch, -1 for an empty bin; may be NULL in forward.  Backward
 * OVERWRITES bottom_diff (gather form, roi_pooling_kernel.cu:147-201). */
int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                          const int height, const int width, const int channels,
                          const int pooled_height, const int pooled_width,
                          const float* bottom_rois, float* top_data, int* argmax_data,
                          cudaStream_t stream);
int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                           const int num_rois, const int height, const int width,
                           const int channels, const int pooled_height, const int pooled_width,
                           const float* bottom_rois, float* bottom_diff, const int* argmax_data,
                           cudaStream_t stream);

/* ------------------------------------------------------- similarity + losses (DVSA) ---- */

/* Replaces DVSA.forward -- model.py:517-614 -- and the backward autograd derives from it
 * (model.py:772), as one fused forward kernel and one fused backward kernel.
 *   vis_feats  (Na*Ns*Nb, D)   region embeddings, rows ordered (segment, frame, box)
 *   word_feats (Na*Ne, D)      query embeddings, rows ordered (segment, padded entity slot)
 *   entities_length (Na) int32 DEVICE: number of real queries per segment (0..Ne)
 *   D_ind (Na*Ns, Na*Ne) int64, D_sim (Na*Ns, Na*Ne) f32: argmax / max over the Nb boxes of the
 *     masked similarity (model.py:608-612; first maximal index on ties)
 *   margin_loss (1) f32: 10*(mean(frame_score) + vis_lam*vis_loss) when train != 0, else
 *     10*mean(frame_score)  (model.py:606)
 * The workspace carries the saved state from forward to backward and must not be touched in
 * between; it must be ZERO-FILLED ONCE after allocation (the kernels leave their arrival
 * counters zeroed on exit).  grad_margin_loss (1) f32 DEVICE = dL/d(margin_loss); NULL = the step
 * wrapper's L1Loss(margin_loss, 0).backward() (model.py:771-772): sign(margin_loss), which the
 * forward left in the workspace.  entities_length values above Ne select all Ne columns and still
 * divide by the raw length, like the reference's slicing (model.py:535-538).
 * grad_vis (Na*Ns*Nb, D) and grad_word (Na*Ne, D) are fully written. */
size_t nafae_ground_workspace_bytes(int Na, int Ns, int Nb, int Ne, int D);
int nafae_ground_forward(const float* vis_feats, const float* word_feats,
                         const int* entities_length, int Na, int Ns, int Nb, int Ne, int D,
                         float Delta, float vis_lam, int train, int64_t* D_ind, float* D_sim,
                         float* margin_loss, void* workspace, size_t workspace_bytes,
                         cudaStream_t stream);
/* `groups` INDEPENDENT batches in one launch (gridDim.y): group g reads vis_feats + g*Na*Ns*Nb*D,
 * word_feats + g*Na*Ne*D, entities_length + g*Na, and writes D_ind / D_sim + g*(Na*Ns)*(Na*Ne),
 * margin_loss[g], using workspace + g*nafae_ground_workspace_bytes(...).  The evaluation sweep runs
 * the reference's batch_size_val = 1 (model.py:514,804) for many segments at once this way: every
 * segment sees only its own queries, exactly as Na = 1 calls would. */
int nafae_ground_forward_batched(const float* vis_feats, const float* word_feats,
                                 const int* entities_length, int groups, int Na, int Ns, int Nb, int Ne,
                                 int D, float Delta, float vis_lam, int train, int64_t* D_ind,
                                 float* D_sim, float* margin_loss, void* workspace,
                                 size_t workspace_bytes, cudaStream_t stream);
/* Same contract and results as nafae_ground_forward, with the regions x queries x dim contraction
 * (S_ = vis_feats @ word_feats^T, model.py:548) on the tcgen05 tensor cores: 128-row tiles, fp32
 * operands fed as three tf32 products per K step (hi*hi + lo*hi + hi*lo), fp32 accumulation in tensor
 * memory; the argmax over boxes is re-checked in exact fp32 whenever the two best candidates are
 * within 6e-5, so D_ind is the fp32 pick.  Bound for this path: D_sim / margin_loss within 2e-4
 * relative (+1e-5 absolute) of the fp32 path.  Nb <= 128.  Two launches (tiles, then the per-segment
 * phases).  Which form is faster depends on the shape -- profiles/RESULTS.md has the A/B. */
int nafae_ground_forward_tc(const float* vis_feats, const float* word_feats, const int* entities_length,
                            int Na, int Ns, int Nb, int Ne, int D, float Delta, float vis_lam, int train,
                            int64_t* D_ind, float* D_sim, float* margin_loss, void* workspace,
                            size_t workspace_bytes, cudaStream_t stream);
int nafae_ground_backward(const float* grad_margin_loss, const float* vis_feats,
                          const float* word_feats, const int* entities_length, int Na, int Ns,
                          int Nb, int Ne, int D, float Delta, float vis_lam, int train,
                          const int64_t* D_ind, const float* D_sim, float* grad_vis,
                          float* grad_word, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream);

/* Device-side postprocess -- model.py:457-474: keeps the a'==a diagonal blocks of D_ind / D_sim
 * and turns the box index into a global row a*Ns*Nb + s*Nb + b.
 * out_ind (Na, Ns, Ne) int64, out_sim (Na, Ns, Ne) f32. */
int nafae_ground_postprocess(const int64_t* D_ind, const float* D_sim, int Na, int Ns, int Nb,
                             int Ne, int64_t* out_ind, float* out_sim, cudaStream_t stream);

/* Evaluation sweep on the device: postprocess (model.py:457-474) + record_det (model.py:477-487) +
 * box accuracy (lib/datasets/youcook_eval.py:241-336) for `num_segments` independent segments whose
 * D_ind / D_sim are in the batched Na = 1 layout (num_segments, Ns, Ne) of
 * nafae_ground_forward_batched.  rois (num_segments*Ns*Nb, 5) as produced by nafae_proposal_tail.
 * Per slot (segment, frame, entity): out_image_ids = image_id_base + segment*Ns + frame (or -1 for a
 * padded entity slot e >= entities_length[segment]), out_box_rows = the global box row, out_boxes
 * (.., 4), out_confs; any of the four may be NULL.  With gt_boxes (num_segments, Ns, Ne, 4) f64 -- ONE
 * ground-truth box per (image, label), the well-formed annotation case -- and gt_classes
 * (num_segments, Ne) int32, every real slot adds 1 to class_count[class] and, when
 * overlap >= gt_thr (the reference's +1 pixel convention and NumPy dtypes), to class_match[class]:
 * the class_match_count / class_count vectors box_accuracy reduces to macro / micro accuracy.  The
 * counters accumulate across calls (zero them once). */
int nafae_eval_record(const int64_t* D_ind, const float* D_sim, const int* entities_length,
                      const float* rois, int num_segments, int Ns, int Nb, int Ne, int64_t image_id_base,
                      int64_t* out_image_ids, int64_t* out_box_rows, float* out_boxes, float* out_confs,
                      const double* gt_boxes, const int* gt_classes, float gt_thr, int num_classes,
                      int* class_match, int* class_count, cudaStream_t stream);

/* ------------------------------------------------------- bridge: tensor-core GEMM ---- */

/* First slice of the "bridge" between RoIAlign and the scoring head: the frozen, inference-only fully
 * connected layers of RCNN_top -- lib/model/faster_rcnn/vgg16_rpn.py:35,56-61 (VGG16 fc6 / fc7:
 * Linear + ReLU; the Dropout between them is a no-op because the detector runs in eval mode,
 * model.py:651,673) -- as one tcgen05 (5th-generation tensor core) kernel per layer:
 *     C[M, N] = act(A[M, K] . B[N, K]^T + bias[N])
 * A (M, K) bf16 row-major = the pooled RoI features viewed as (R, C*7*7) (vgg16_rpn.py:58); B (N, K)
 * bf16 row-major = nn.Linear.weight as PyTorch stores it; bias (N) fp32 or NULL; C (M, N) fp32, or bf16
 * with NAFAE_GEMM_OUT_BF16 (feeds the next layer).  fp32 accumulation in tensor memory.  K % 8 == 0,
 * 16-byte aligned buffers.  bf16 inputs: expect ~1e-2 relative agreement with the fp32 layer. */
#define NAFAE_GEMM_RELU 1u
#define NAFAE_GEMM_OUT_BF16 2u
int nafae_gemm_bf16_tn(const void* A, const void* B, const float* bias, void* C, int M, int N, int K,
                       unsigned flags, cudaStream_t stream);

/* ------------------------------------------------------- step wrapper: clip + Adam ---- */

/* Replaces clip_grad_norm_(ground_model.parameters(), args.clip) + optimizer.step() --
 * model.py:773-774 (Adam, lr 1e-3, weight_decay 1e-5: model.py:1030-1036) -- over ONE flat fp32
 * buffer of the trainable parameters (the bucket the data-parallel all-reduce averages), two launches,
 * no host synchronisation.  grad is scaled in place by min(1, max_norm / (||grad||_2 + 1e-6)) like
 * clip_grad_norm_ (max_norm <= 0 disables clipping), then torch.optim.Adam's update (L2 weight decay
 * added to the gradient, bias correction by the step count kept in the workspace).  The norm is summed
 * in a fixed order: replicas that start identical and see the same averaged gradient stay identical.
 * workspace: nafae_clip_adam_workspace_bytes() bytes, ZERO-FILLED ONCE (holds the step count). */
size_t nafae_clip_adam_workspace_bytes(void);
int nafae_clip_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, size_t n,
                         float lr, float beta1, float beta2, float eps, float weight_decay,
                         float max_norm, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* ------------------------------------------------- data-parallel gradient all-reduce ---- */

/* New functionality (the reference is single-GPU: --mGPUs is parsed and never read, model.py:91-99).
 * Two-shot all-reduce (AVG) of a flat fp32 gradient bucket over NVLink peer memory, one process per
 * GPU, world <= 8.  Each rank allocates a symmetric buffer with nafae_ar_alloc (cudaMalloc +
 * cudaIpcGetMemHandle), exchanges the 64-byte handles out of band (torch.distributed / MPI), maps the
 * peers' buffers with nafae_ar_open, and writes its gradients at byte offset nafae_ar_data_offset().
 * nafae_allreduce_avg launches ONE kernel (no host state, graph-capturable) on `stream`; all ranks
 * must launch it the same number of times with the same count / num_ctas.  Summation order is fixed
 * (rank 0..world-1), so every replica holds bit-identical averages afterwards.
 * bufs: host array of `world` device pointers as mapped in this process, bufs[rank] = own buffer.
 * count_floats must be a multiple of 4*world (nafae_ar_buffer_bytes pads).
 * cta_threads: 0 = bulk-copy kernel (cp.async.bulk pulls the slice from every rank into shared
 * memory, reduces, and pushes the result into every rank's buffer; one CTA per SM, num_ctas = the SMs
 * nafae_set_reserved_sms keeps free); 256 / 128 = per-thread 16-byte loads (four / eight CTAs per
 * SM).  flags: NAFAE_AR_VARIANT(v) selects the bulk-copy kernel's ring shape (0 = 4 slots, one
 * issuing thread; 1 = 3 slots of twice the chunk size, one issuing lane per peer);
 * NAFAE_AR_WIDTH(w) forces the compile-time world bound (2, 4 or 8, >= world; tests only).
 * When it overlaps the RoIAlign kernel, enqueue nafae_gate_wait first (see there).
 * Cross-GPU waits are bounded (2 s): a dead peer sets a sticky error word instead of hanging. */
#define NAFAE_AR_VARIANT(v) ((unsigned)(v) & 0xfu)
#define NAFAE_AR_WIDTH(w) (((unsigned)(w) & 0xffu) << 8)
size_t nafae_ar_buffer_bytes(size_t count_floats, int world);
size_t nafae_ar_data_offset(void);
int nafae_ar_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64);
int nafae_ar_open(const unsigned char* handle64, void** peer_ptr);
int nafae_ar_close(void* peer_ptr);
int nafae_ar_free(void* dev_ptr);
int nafae_allreduce_avg(void* const* bufs, int rank, int world, size_t count_floats, int num_ctas,
                        int cta_threads, unsigned flags, cudaStream_t stream);

/* NVLS form of the same collective: the bucket lives in memory bound to an NVSwitch MULTICAST object,
 * the sum is formed INSIDE the switch (multimem.ld_reduce) and the result is replicated to every rank
 * by one multicast store (multimem.st) -- per rank the SMs move count/world floats each way instead of
 * (world-1)/world of the bucket twice.  Setup (once, host side, one process per GPU):
 *   root : nafae_mc_create(world, bytes, &h, &fd)   -> hand `fd` to the other processes (SCM_RIGHTS)
 *   other: nafae_mc_import(fd, world, bytes, &h)
 *   all  : nafae_mc_add_device(h); <host barrier>; nafae_mc_bind(h, &uc, &mc); <host barrier>
 * `uc` = this rank's own copy (zero-filled; gradients go at uc + nafae_ar_data_offset()), `mc` = the
 * multicast view.  bytes = nafae_mc_buffer_bytes(count, world).  nafae_mc_supported() = 1 when the
 * current device can do this (CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED); otherwise use the peer-memory
 * kernel above.  nafae_allreduce_mc launches ONE kernel (graph-capturable, same calling rules as
 * nafae_allreduce_avg); cta_threads 0 (= 512), 256, 512 or 1024.  One rank reduces each element, so
 * replicas are bit-identical; the order of the in-switch sum is the hardware's.
 * nafae_allreduce_mc_error (host-synchronising) reports whether a cross-GPU wait ever timed out. */
int nafae_mc_supported(void);
size_t nafae_mc_buffer_bytes(size_t count_floats, int world);
int nafae_mc_create(int world, size_t bytes, void** handle, int* fd_out);
int nafae_mc_import(int fd, int world, size_t bytes, void** handle);
int nafae_mc_add_device(void* handle);
int nafae_mc_bind(void* handle, void** uc_ptr, void** mc_ptr);
size_t nafae_mc_size(void* handle);
int nafae_mc_free(void* handle);
int nafae_allreduce_mc(void* uc_base, void* mc_base, int rank, int world, size_t count_floats,
                       int num_ctas, int cta_threads, cudaStream_t stream);
int nafae_allreduce_mc_error(void* uc_base);

#ifdef __cplusplus
}
#endif

#endif /* NAFAE_B200_H_ */
