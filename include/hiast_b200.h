/*
 * hiast_b200 -- C ABI of the B200-native (sm_100a) post-logit self-training hot path of HIAST.
 *
 * The reference (bupt-ai-cz/HIAST) is pure Python and has no FFI; its extension point is the
 * string-keyed registry (code/utils/registry/registry.py:6-43).  Each entry point below replaces
 * the torch/numpy op chain of one reference function (cited per function, paths relative to
 * /root/reference/code) and is what a ctypes / cffi binding on the reference side calls
 * (see INTEGRATION.md).  The Python classes in hiast_b200/ mirror the reference's interface
 * on top of these calls.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; buffers are caller-owned;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *     asynchronous on that stream and re-entrant per stream; nothing is allocated internally
 *     (workspace sizes come from the *_bytes query functions);
 *   - return value: HIAST_OK or a negative HIAST_ERR_* code, never an exception;
 *   - tensors are dense, row-major, in the reference's layouts (logits NCHW float32);
 *   - labels written by the library are uint8 with 255 = ignore (the PNG payload of
 *     pseudo_label_generator.py:46); label inputs may be uint8 or int64 (`*_bytes` = 1 or 8).
 */
#ifndef HIAST_B200_H_
#define HIAST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HIAST_API __attribute__((visibility("default")))
#else
#define HIAST_API
#endif

#define HIAST_OK                 0
#define HIAST_ERR_INVALID_ARG   -1
#define HIAST_ERR_UNSUPPORTED   -2
#define HIAST_ERR_CUDA          -3
#define HIAST_ERR_WORKSPACE     -4
#define HIAST_ERR_IO            -5

#define HIAST_IGNORE_LABEL      255
#define HIAST_KEY_ONE           0x3C00  /* fp16 bit pattern of 1.0: the largest histogram key      */
#define HIAST_MAX_CLASSES       255

/* region selector of the consistency loss (losses.py:75-84) */
#define HIAST_REGION_IGNORED    0
#define HIAST_REGION_CONFIDENT  1
#define HIAST_REGION_ALL        2

/* loss term bit mask */
#define HIAST_TERM_CE           1
#define HIAST_TERM_KLD          2
#define HIAST_TERM_ENT          4
#define HIAST_TERM_CST          8
/* kind of the consistency term, OR-ed into `terms` (default 0 = SoftCE, losses.py:39-61) */
#define HIAST_CST_SOFTCE         0   /* -logp_c * t_c,            t = soft targets in [0,1]                           */
#define HIAST_CST_KLDIV         16   /* xlogy(tp,tp) - tp*logp_c, t = target LOGITS, tp = softmax(t), losses.py:16-23 */
#define HIAST_CST_MSE           32   /* (z_c - t_c)^2,                                                losses.py:9-13  */
#define HIAST_CST_SOFTCE_LOGITS 48   /* SoftCE with t = teacher LOGITS: the trainer's F.softmax (consistency_self_training_trainer.py:119) fused in */

/* ---- library ---------------------------------------------------------------------------- */
HIAST_API int         hiast_version(void);                 /* 1000*major + minor                          */
HIAST_API const char* hiast_status_string(int status);
HIAST_API int         hiast_last_cuda_error(void);         /* cudaError_t of the last HIAST_ERR_CUDA      */
HIAST_API int         hiast_device_sm_count(void);         /* SMs of the current device (0 if no device)  */

/* ---- (1) instance-adaptive selector ----------------------------------------------------- */

/* Smallest fp16 key a confidence 1/sum(exp) can round to with C classes: fp16_rn(1.0f/C).  */
HIAST_API int    hiast_ias_key_lo(int C);
/* Histogram buffers are uint32 [n_groups][C][row_stride]; a row holds the HIAST_KEY_ONE - key_lo + 1
 * bins of one (group, class), padded to a multiple of 4 words so that rows are 16-byte aligned.  */
HIAST_API int    hiast_ias_hist_row_stride(int key_lo);
HIAST_API size_t hiast_ias_hist_bytes(int n_groups, int C, int key_lo);

/* a1+a2  workflows/pseudo_label_generator.py:192-193,198-201
 * softmax over C + first-index max (bit-exact with ATen's CUDA softmax -> max(dim=1)), fused
 * with the per-(group, class) histogram of fp16_rn(conf) bit patterns.  Image i belongs to
 * group i / group_size (the reference's DataLoader batch).  `accumulate` = 0 zeroes `hist`
 * first.  `hist_mode`: bits 0-7 select the histogram strategy (0 = library default); bits 8-15 = number of SMs the
 * default kernel leaves WITHOUT one of its CTAs (its CTAs take whole SMs; the free SMs are where a caller runs the
 * threshold scan and the mask pass of earlier windows concurrently on another stream).
 *   logits f32 [n_images,C,H,W];  conf f32 [n_images,H,W];  label u8 [n_images,H,W]          */
HIAST_API int hiast_ias_softmax_hist(const float* logits, int n_images, int C, int H, int W,
                           int group_size, int key_lo, int accumulate, int hist_mode,
                           float* conf, uint8_t* label, uint32_t* hist, void* stream);

/* a1+a2 with the bilinear up-sampling of the step before the path fused in (SURVEY.md 8f rank 1;
 * sseg/models/segmentors/self_training_segmentor.py:27: F.interpolate(mode='bilinear', align_corners=True)).
 * logits_lr f32 [n_images,C,h_in,w_in] is the network output at its own resolution; conf / label / hist are
 * those of softmax(interpolate(logits_lr, (H,W))).max(1), bit-identical to ATen's CUDA path, without the
 * full-resolution tensor ever existing.  C in {19,16}, W % 4 == 0, up-sampling only; else HIAST_ERR_UNSUPPORTED. */
HIAST_API int hiast_ias_upsample_softmax_hist(const float* logits_lr, int n_images, int C, int h_in, int w_in,
                                    int H, int W, int group_size, int key_lo, int accumulate,
                                    float* conf, uint8_t* label, uint32_t* hist, void* stream);

/* a2 alone, for callers that already hold conf/label (the signature of
 * select_and_save_confident_label / get_ias_threshold takes them, :67,:171).
 * label_bytes = 1 (uint8) or 8 (int64); label_u8_out (nullable) receives the uint8 copy.     */
HIAST_API int hiast_ias_conf_hist(const float* conf, const void* label, int label_bytes,
                        int n_images, int64_t HW, int C, int group_size, int key_lo,
                        int accumulate, uint8_t* label_u8_out, uint32_t* hist, void* stream);

/* a3+a4  :171-179, :207-209
 * Sequential per-class EMA threshold scan over groups.  Step g, class c:
 *   q = 1 - alpha * thr^gamma;  temp = float32(quantile_linear({keys of (g,c)} U {thr}, q));
 *   thr = beta*thr + double(float(1-beta) * temp);  thr >= 1 -> 0.999
 * `hist` is converted IN PLACE to inclusive prefix sums.  thr_state f64[C] is read and
 * updated; thr_groups f64 [n_groups,C] receives the threshold in force for each group (the
 * post-update value, which is the one the reference masks that same batch with, :211);
 * temp_groups f32 [n_groups,C] (nullable) receives the quantiles.  *error_flag (device int,
 * nullable) is OR-ed with: 1 if some q left [0,1] (numpy raises ValueError there); 2 if some
 * float32 quantile is not certified independent of the last-bit rounding of the host libm's
 * pow() (see hiast_b200/csrc/scan_math.h; bit-exactness versus numpy is certified when 0).   */
HIAST_API int hiast_ias_threshold_scan(uint32_t* hist, int n_groups, int C, int key_lo,
                             double alpha, double beta, double gamma,
                             double* thr_state, double* thr_groups, float* temp_groups,
                             int* error_flag, void* stream);

/* a3+a4 with the multi-GPU hand-off of the threshold state fused in (the only cross-GPU dependency of the path
 * is this f64[C] state; the reference is single-process).  Token ring over peer memory: a MAILBOX is hiast_ring_mailbox_bytes()
 * of device memory created by hiast_ring_create, which also returns a 64-byte CUDA IPC handle that another process on the node
 * turns into a peer pointer with hiast_ring_open.  The scan kernel of window w
 *   - token_in  != NULL: CTA c waits until slot c of THIS GPU's mailbox carries a sequence number >= in_seq and starts from
 *                        the threshold stored there instead of thr_state[c];
 *   - token_out != NULL: stores its final threshold into slot c of the NEXT GPU's mailbox (peer pointer: the store travels
 *                        over NVLink) and releases out_seq there (st.release.sys).
 * No receive / send kernel, no host involvement per hop.  A wait of more than 10 s sets bit 8 of *error_flag instead of
 * hanging.  With both tokens NULL this is hiast_ias_threshold_scan.  Sequence numbers must grow along the ring.          */
HIAST_API size_t hiast_ring_mailbox_bytes(void);
HIAST_API int hiast_ring_create(void** local_box_out, void* ipc_handle_out /* 64 bytes, host */);
HIAST_API int hiast_ring_open(const void* ipc_handle /* 64 bytes, host */, void** peer_box_out);
HIAST_API int hiast_ring_close(void* peer_box);
HIAST_API int hiast_ring_destroy(void* local_box);
HIAST_API int hiast_ias_threshold_scan_ring(uint32_t* hist, int n_groups, int C, int key_lo,
                                  double alpha, double beta, double gamma,
                                  double* thr_state, double* thr_groups, float* temp_groups, int* error_flag,
                                  const void* token_in, uint64_t in_seq, void* token_out, uint64_t out_seq, void* stream);

/* a5+a6+a7(sums)  :71-89, :96-99
 * plbl = conf < thr[group][label] ? 255 : label (float32 conf compared with the float64
 * threshold); per-image counts of the kept labels; per-group fixed-point (2^-32) sums of the
 * kept confidences per class.  counts / confsum are ACCUMULATED into (zero them first).
 *   thr_groups f64 [n_groups,C]; plbl u8 [n_images,HW]; counts i64 [n_images,C];
 *   confsum u64 [n_groups,C]                                                                  */
HIAST_API int hiast_ias_select(const float* conf, const uint8_t* label, const double* thr_groups,
                     int n_images, int64_t HW, int C, int group_size,
                     uint8_t* plbl, int64_t* counts, uint64_t* confsum, void* stream);

/* a7  :95-105   class_mean_probs EMA over groups (first touch initialises):
 *   m = float32(sum/count);  cmp = (cmp == 0) ? m : cmp*g + double(m * float(1-g))           */
HIAST_API int hiast_ias_meanprob_scan(const uint64_t* confsum, const int64_t* counts,
                            int n_images, int group_size, int n_groups, int C, double cp_gamma,
                            double* mean_state, void* stream);

/* a1-a7 in ONE launch for a window of images on a single GPU (the whole loop body of
 * IASPseudoGenerator.run, :190-211, for n_images / group_size consecutive batches):
 * phase A, the threshold chain and phase C run in one persistent kernel; the conf / label spill stays in L2
 * (conf_scratch f32 [n,HW] / label_scratch u8 [n,HW] hold NO defined values afterwards).  Results are
 * bit-identical to hiast_ias_softmax_hist + hiast_ias_threshold_scan + hiast_ias_select:
 *   thr_state (in/out), thr_groups, temp_groups (may be NULL), plbl (hist is scratch: raw counts on return),
 *   counts i64 [n,C] and confsum u64 [G,C] (both OVERWRITTEN here, not accumulated), *error_flag |= 1 / 2 as in
 *   hiast_ias_threshold_scan, |= 4 if the kernel gave up waiting (internal error).
 * workspace: hiast_ias_fused_workspace_bytes.  flags: bit 0 = keep the spill lines (no discard.global.L2);
 * bits 4..7 = groups in flight (0 = default 2).  Returns HIAST_ERR_UNSUPPORTED for shapes the fused kernel
 * does not cover (C not in {16, 19}, HW % 4 != 0, unaligned buffers): use the three calls above.          */
HIAST_API size_t hiast_ias_fused_workspace_bytes(int n_images, int group_size);
HIAST_API int hiast_ias_fused_window(const float* logits, int n_images, int C, int H, int W, int group_size,
                           int key_lo, double alpha, double beta, double gamma,
                           float* conf_scratch, uint8_t* label_scratch, uint32_t* hist,
                           double* thr_state, double* thr_groups, float* temp_groups,
                           uint8_t* plbl, int64_t* counts, uint64_t* confsum, int* error_flag,
                           void* workspace, size_t workspace_bytes, int flags, void* stream);

/* CBST policy (8f rank 3)  :142-165.  Adds to hist u32 [C][row_stride] (one histogram for the whole data set,
 * accumulated over calls) the fp16 keys of the pixels whose rank among the pixels of their class, in raster
 * order over the images of their batch (group), is a multiple of sample_interval.  workspace: see
 * hiast_cbst_workspace_bytes.  C <= 48.                                                                   */
HIAST_API size_t hiast_cbst_workspace_bytes(int n_images, int64_t HW, int C);
HIAST_API int hiast_cbst_sample_hist(const float* conf, const uint8_t* label, int n_images, int64_t HW, int C,
                           int group_size, int sample_interval, int key_lo, uint32_t* hist,
                           void* workspace, size_t workspace_bytes, void* stream);
/* thr f64[C] = np.quantile(samples_c, q) ('linear') read off that histogram.  *error_flag |= 4 for a class
 * without samples (numpy raises IndexError; thr = NaN), |= 1 for q outside [0,1].                          */
HIAST_API int hiast_cbst_quantile(const uint32_t* hist, int C, int key_lo, double q, double* thr,
                        int* error_flag, void* stream);

/* ---- (2) hard-aware copy-paste  sseg/datasets/preprocessor.py:102-112 ------------------- */
/* For image i with donor d = donor_index ? donor_index[i] : i, per pixel:
 *   M = hard[donor_lbl]; img = M ? donor_img : img; lbl = M ? donor_lbl : lbl;
 *   cp_mask = M ? donor_lbl : cp_mask.   hard_lut_host: 256-bit set (8 x uint32, HOST memory).
 *   img u8 [n,HW,3] (HWC); lbl, cp_mask u8 [n,HW]; donors likewise, indexed by d.            */
HIAST_API int hiast_copy_paste(uint8_t* img, uint8_t* lbl, uint8_t* cp_mask,
                     const uint8_t* donor_img, const uint8_t* donor_lbl,
                     const int32_t* donor_index, int n_images, int64_t HW,
                     const uint32_t* hard_lut_host, void* stream);

/* ---- (3) region-adaptive regularisation + consistency losses ---------------------------- */
/* a10-a15  sseg/models/segmentors/self_training_segmentor.py:30-53,128-163; losses.py:32-89
 * One pass over (z, t, plbl): log-softmax once per pixel, then the four masked reductions
 *   sums[0] = sum_conf -logp[y]            (CE,  / n_conf)
 *   sums[1] = sum_conf sum_c -logp_c / C   (KLD, / (C*n_conf))
 *   sums[2] = sum_ign  -sum_c p_c logp_c   (ENT, / (C*n_ign))
 *   sums[3] = sum_region sum_c -logp_c t_c (CST, / counts[2])
 *   counts  = { n_conf, n_ign, #nonzero fp32 products in the CST region }
 * `terms` = mask of HIAST_TERM_*; t may be NULL without HIAST_TERM_CST.  Deterministic
 * (two-stage reduction through `workspace`).                                                 */
HIAST_API size_t hiast_st_loss_workspace_bytes(int B, int C, int64_t HW);
HIAST_API int hiast_st_loss_fwd(const float* z, const float* t, const void* plbl, int plbl_bytes,
                      int B, int C, int64_t HW, int region, int terms,
                      double* sums, int64_t* counts,
                      void* workspace, size_t workspace_bytes, void* stream);
/* grad_z of  sum_k scales[k] * (unnormalised term k), scales f32[4] on the DEVICE
 * (weight * upstream grad / denominator, computed by the caller without a host sync).       */
HIAST_API int hiast_st_loss_bwd(const float* z, const float* t, const void* plbl, int plbl_bytes,
                      int B, int C, int64_t HW, int region, int terms,
                      const float* scales, float* grad_z, void* stream);

/* a10-a15 forward AND backward in ONE pass over (z, t, plbl): 236 B/px instead of the 160 + 236 of the two calls above.
 * A label-only pre-pass fixes n_conf / n_ign; the main pass accumulates the forward sums and writes
 *   grad_z = sum_k scales_used[k] * d(term k)/dz,   scales_used[k] = float(double(grad_weights[k]) / divisor_k)
 * with divisor = {n_conf, C*n_conf, C*n_ign, C*n_region}: the gradient for ASSUMED upstream gradients grad_weights f32[4]
 * (device; the caller's loss weights times the upstream scalar it expects).  sums / counts as hiast_st_loss_fwd;
 * scales_used f32[4] (device) records the assumption.  When autograd later delivers the real upstream gradients, the
 * caller computes the scales they imply and calls hiast_st_loss_bwd_checked: if they are the same bits as scales_used the
 * kernel exits at once (grad_z is already right), else it rewrites grad_z like hiast_st_loss_bwd -- exact in every case
 * (changed loss scale, SoftCE divisor != C*n_region because some product was exactly 0).  SoftCE kind, C in {16, 19},
 * even HW only: otherwise HIAST_ERR_UNSUPPORTED and nothing is launched (use the two calls above).                    */
HIAST_API size_t hiast_st_loss_fused_workspace_bytes(int B, int C, int64_t HW);
HIAST_API int hiast_st_loss_fused(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C, int64_t HW,
                        int region, int terms, const float* grad_weights, double* sums, int64_t* counts,
                        float* scales_used, float* grad_z, void* workspace, size_t workspace_bytes, void* stream);
HIAST_API int hiast_st_loss_bwd_checked(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                              int64_t HW, int region, int terms, const float* scales, const float* scales_used,
                              float* grad_z, void* stream);

/* The same two calls with the arithmetic AROUND the kernels done on the device as well (self_training_segmentor.py:30-53 divides
 * every sum by its count and autograd multiplies the upstream gradients back in: a dozen tiny launches per step otherwise).
 * hiast_st_loss_fused_terms additionally writes losses f32[4] = float(sums[k] / divisors[k]) (0 for a disabled term; NaN for an
 * empty region, the reference's 0/0) and divisors f64[4] = {n_conf, C*n_conf, C*n_ign, counts[2]} (both or neither may be NULL).
 * hiast_st_loss_bwd_checked_terms takes the upstream gradient of each of the four loss terms as a device pointer to one f32
 * (NULL: no gradient flows into that term) and derives scales[k] = float(double(*gout_k) / divisors[k]) itself; if upstream_out
 * != NULL it also stores gout_{k0} / hint_weights[k0] there (the upstream scalar the next forward call should assume).
 * term_weights f32[4] (device, may be NULL): the caller's loss weights.  With them `losses` holds the WEIGHTED terms
 * w_k * loss_k (float32 product, what `weight * loss` gives in torch) and the backward call takes the upstream gradients of
 * those weighted terms: gout_k = *gout_k_ptr * w_k -- the reference's `w * loss` products and their MulBackward nodes
 * (self_training_segmentor.py:37-52) without their eight launches.                                                         */
HIAST_API int hiast_st_loss_fused_terms(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C, int64_t HW,
                              int region, int terms, const float* grad_weights, double* sums, int64_t* counts,
                              float* scales_used, float* grad_z, float* losses, double* divisors,
                              const float* term_weights, void* workspace, size_t workspace_bytes, void* stream);
HIAST_API int hiast_st_loss_bwd_checked_terms(const float* z, const float* t, const void* plbl, int plbl_bytes, int B, int C,
                                    int64_t HW, int region, int terms, const float* gout_ce, const float* gout_kld,
                                    const float* gout_ent, const float* gout_cst, const double* divisors,
                                    const float* scales_used, float* grad_z, const float* term_weights,
                                    const float* hint_weights, int k0, float* upstream_out, void* stream);

/* ---- (4) confusion matrix / mIoU  utils/metrics.py:6-19 --------------------------------- */
/* cm i64 [(K+1),(K+1)] (rows = target, cols = pred, index K = value outside [0,K)),
 * ACCUMULATED over the pixels whose target != ignore_index.  If pred_masked_out != NULL it
 * receives pred with the ignored pixels overwritten (the reference's in-place side effect,
 * metrics.py:12); it may alias pred.  elem_bytes = 1 or 8 for both arrays.                   */
HIAST_API int hiast_confusion_matrix(const void* pred, const void* target, int elem_bytes, int64_t n,
                           int K, int ignore_index, void* pred_masked_out,
                           int64_t* cm, void* stream);
/* Same with pred = first-index argmax over C of logits [B,C,HW] (base_trainer.py:173).       */
HIAST_API int hiast_confusion_from_logits(const float* logits, const void* target, int target_bytes,
                                int B, int C, int64_t HW, int K, int ignore_index,
                                int64_t* cm, void* stream);
/* area_intersection, area_union f32 [K] from cm (diag; row+col-diag).                        */
HIAST_API int hiast_iou_from_confusion(const int64_t* cm, int K, float* intersection, float* area_union,
                             void* stream);

/* ---- EMA teacher update (8f rank 4)  utils/utils.py:115-123 ------------------------------ */
/* param_k = param_k * gamma + param_q * (1 - gamma) for a whole model in ONE launch (three float32 roundings per
 * element like the reference; gamma / one_minus_gamma are the float32 roundings of the Python floats gamma and
 * 1 - gamma).  segs_dev: device table of n_seg x {void* k, const void* q, int64 n}; the segments are cut into chunks
 * of chunk_elems elements (multiple of 4): chunk_seg_dev i32 [n_chunks] names the segment of each chunk, chunk_off_dev
 * i64 [n_chunks] its first element.  All pointers are float32 device memory.                       */
HIAST_API int hiast_ema_update(const void* segs_dev, const int32_t* chunk_seg_dev, const int64_t* chunk_off_dev,
                     int n_chunks, int chunk_elems, float gamma, float one_minus_gamma, void* stream);
/* buffer_k = buffer_q (:120-121) with the same tables; n and the chunk size are in BYTES (multiple of 16).       */
HIAST_API int hiast_multi_copy(const void* segs_dev, const int32_t* chunk_seg_dev, const int64_t* chunk_off_dev,
                     int n_chunks, int chunk_bytes, void* stream);

/* ---- pseudo-label PNG writer (8f rank 2)  workflows/pseudo_label_generator.py:43-46 -------- */
/* Replaces `cv2.imwrite(path, plbl.astype(np.uint8))`: encodes n_images uint8 label maps [n,H,W] into complete
 * 8-bit gray PNG files (read back by base_dataset.py:158-170 `Image.open`), packed back to back in `out`:
 * file i = out[offsets[i] .. offsets[i+1]).  offsets is int64 [n_images + 1] on the device and always written;
 * if offsets[n_images] > out_capacity nothing else is written (call again with a larger buffer;
 * n_images * hiast_png_max_bytes(H, W) always suffices).  `out` must be 4-byte aligned, the workspace 256-byte
 * aligned.  Format (Up filter, fixed-Huffman DEFLATE with distance-1 matches, one IDAT chunk per segment, stored
 * fallback, Adler-32 and CRC-32 on the device): oracle/png.py restates it byte for byte.  W <= 32768.          */
HIAST_API size_t hiast_png_workspace_bytes(int n_images, int H, int W);
HIAST_API size_t hiast_png_max_bytes(int H, int W);
HIAST_API int hiast_png_segments(int H, int W);
HIAST_API int hiast_png_encode(const uint8_t* labels, int n_images, int H, int W, uint8_t* out, size_t out_capacity,
                     int64_t* offsets, void* workspace, size_t workspace_bytes, void* stream);
/* HOST function: writes file i = blob_host[offsets_host[i] .. offsets_host[i+1]) to paths_host[i] (create / truncate,
 * mode 0644) for i < n_files with n_threads POSIX writer threads; *errno_out = first errno (0 if none).  All pointers are
 * host memory (the blob is the pinned copy of hiast_png_encode's output).  Returns HIAST_ERR_IO if any file failed.     */
HIAST_API int hiast_write_files(const char* const* paths_host, const uint8_t* blob_host, const int64_t* offsets_host,
                      int n_files, int n_threads, int* errno_out);

/* ---- pseudo-label reader side (8f rank 2)  sseg/datasets/loader/base_dataset.py:176 --------- */
/* `cv2.resize(lbl, (Wd, Hd), interpolation=cv2.INTER_NEAREST)` for n uint8 label maps [n,Hs,Ws] -> [n,Hd,Wd]:
 * sx = min(floor(x * inv_scale_x), Ws - 1), sy likewise, in double; the caller passes OpenCV's own factors
 * inv_scale_x = 1.0 / ((double)Wd / Ws), inv_scale_y = 1.0 / ((double)Hd / Hs).                              */
HIAST_API int hiast_resize_nearest_u8(const uint8_t* src, int n_images, int Hs, int Ws, uint8_t* dst, int Hd, int Wd,
                            double inv_scale_x, double inv_scale_y, void* stream);

/* ---- validator: multi-scale / flip softmax sum + arg-max (8f rank 4)  workflows/validator.py:34-55,92-93 ---- */
/* probs = softmax(logits, dim=1) [+ flip_x(softmax(logits_of_flipped, dim=1))] for float32 [B,C,h,w] tensors
 * (validator.py:37,46,48-50: `pred_result += torch.flip(flip_logits, dims=[3])`); logits_of_flipped may be NULL.   */
HIAST_API int hiast_softmax_flip_sum(const float* logits, const float* logits_of_flipped, int B, int C, int h, int w,
                           float* probs, void* stream);
/* label u8 [B,H,W] = first-index argmax over C of sum_s interpolate(probs_s [B,C,h_s,w_s], (H,W), bilinear,
 * align_corners=True), summed in list order (validator.py:52-55,93).  probs_host / h_host / w_host are HOST arrays of
 * n_scales (<= 8) device pointers / sizes.                                                                     */
HIAST_API int hiast_probs_upsample_argmax(const float* const* probs_host, const int* h_host, const int* w_host, int n_scales,
                                int B, int C, int H, int W, uint8_t* label, void* stream);

/* ---- CE with class weights / refer_labels (8f rank 3)  sseg/models/modules/losses.py:32-36,68-89 ---------------- */
/* refer_labels == NULL: sums[0] = sum_{y != ignore} w[y] * nll, sums[1] = sum_{y != ignore} w[y]   (loss = sums[0] / sums[1],
 *   nn.CrossEntropyLoss(ignore_index, weight)); class_weights f32 [C] or NULL (all ones).
 * refer_labels != NULL: L = CE(weight, reduction='none') [B,H,W], mask = region(refer_labels) [B,1,H,W]; the product
 *   broadcasts to [B,B,H,W] (losses.py:86-87): sums[0] = its sum, count[0] = its non-zero count (loss = sums[0] / count[0]);
 *   labels outside [0,C) contribute 0 (the reference's CrossEntropyLoss would raise on them).
 * Backward: grad = *scale * d(sums[0])/d(logits), scale a device float (g / sums[1] or g / count[0]).                  */
HIAST_API size_t hiast_ce_general_workspace_bytes(int64_t HW);
HIAST_API int hiast_ce_general_fwd(const float* logits, const void* labels, int label_bytes, const float* class_weights,
                         const void* refer_labels, int refer_bytes, int region, int ignore_index, int B, int C,
                         int64_t HW, double* sums, int64_t* count, void* workspace, size_t workspace_bytes, void* stream);
HIAST_API int hiast_ce_general_bwd(const float* logits, const void* labels, int label_bytes, const float* class_weights,
                         const void* refer_labels, int refer_bytes, int region, int ignore_index, int B, int C,
                         int64_t HW, const float* scale, float* grad_logits, void* stream);

/* ---- host side of the pseudo-labelling loop  workflows/pseudo_label_generator.py:189-211 ------------------------
 * HOST functions (they enqueue CUDA work but are not kernels).  What the reference does per batch with
 * `data['images'].cuda()` (:190), per image with cv2.imwrite (:43-46) and per batch with numpy bookkeeping (:82-105)
 * becomes a handful of foreign calls per WINDOW of batches; the interpreter never waits for the GPU.              */

/* H2D staging ring.  Slot s is a caller-owned device buffer; two events per slot order the copy stream against the
 * consumer stream: push = [copy_stream waits until the slot's previous contents were consumed] cudaMemcpyAsync on
 * copy_stream, [consumer_stream waits for the copy]; release = "everything queued on consumer_stream so far has
 * consumed slots first .. first+n-1".  A slot must be released before it is pushed again.                          */
HIAST_API int hiast_stager_create(int n_slots, void** handle_out);
HIAST_API int hiast_stager_destroy(void* handle);
HIAST_API int hiast_stager_push(void* handle, int slot, void* dst_device, const void* src_host, size_t nbytes,
                      void* copy_stream, void* consumer_stream);
HIAST_API int hiast_stager_release(void* handle, int first_slot, int n_slots, void* consumer_stream);

/* Everything a window of n_images owes the host after its thresholds are known, queued by ONE call on `stream`:
 * zero counts / confsum, hiast_ias_select (:71-89), optionally hiast_ias_meanprob_scan (:95-105; mean_state NULL =
 * skip), then either hiast_png_encode (:43-46; blob_dev != NULL) with the copies of the offset table and of the first
 * blob_copy_bytes of the blob to pinned host memory, or (blob_dev NULL, plbl_host != NULL) the copy of the uint8
 * label maps themselves; and the copies of counts / confsum / thr_groups (each nullable).  The kernels run on `stream`;
 * the device-to-host copies run on `copy_stream` behind them (NULL or == stream: on `stream` itself), so that they
 * overlap the next window's kernels -- the caller must make `stream` wait for `copy_stream` before it overwrites the
 * window's device buffers again.  Host pointers must stay valid and untouched until copy_stream reaches this point
 * (use an event or hiast_writer_submit's ticket, both recorded on copy_stream).                                   */
typedef struct HiastWindowEmit {
  const float*   conf;            /* f32 [n,H,W]   phase A output                                   */
  const uint8_t* label;           /* u8  [n,H,W]                                                    */
  const double*  thr_groups;      /* f64 [g,C]     thresholds in force per group (phase B output)   */
  uint8_t*       plbl;            /* u8  [n,H,W]   out                                              */
  int64_t*       counts;          /* i64 [n,C]     out (zeroed here)                                */
  uint64_t*      confsum;         /* u64 [g,C]     out (zeroed here)                                */
  double*        mean_state;      /* f64 [C]       in/out, nullable                                 */
  uint8_t*       blob_dev;        /* PNG blob on the device, nullable                               */
  int64_t*       offsets_dev;     /* i64 [n+1]                                                      */
  void*          png_ws;          /* hiast_png_workspace_bytes(n, H, W)                             */
  uint8_t*       blob_host;       /* pinned                                                         */
  int64_t*       offsets_host;    /* pinned i64 [n+1]                                               */
  uint8_t*       plbl_host;       /* pinned u8 [n,H,W], used when blob_dev == NULL                  */
  int64_t*       counts_host;     /* pinned, nullable                                               */
  uint64_t*      confsum_host;    /* pinned, nullable                                               */
  double*        thr_groups_host; /* pinned, nullable                                               */
  size_t         blob_capacity;   /* bytes of blob_dev                                              */
  size_t         png_ws_bytes;
  size_t         blob_copy_bytes; /* predicted size of the window's files: copied with the window   */
  double         cp_gamma;
  int32_t        n_images, H, W, C, group_size, reserved;
} HiastWindowEmit;
HIAST_API int hiast_ias_emit_window(const HiastWindowEmit* args, void* stream, void* copy_stream);

/* Asynchronous file writer (:43-46 without the interpreter).  submit records an event on `stream` (behind
 * hiast_ias_emit_window's copies) and returns a ticket > 0 at once; a dispatcher thread sleeps on the event, then
 * reads offsets_host[n_files] = the true size of the window's files, fetches blob_dev[bytes_copied .. total) if the
 * predicted copy was short (blob_host_capacity must cover it) and the pool's POSIX writer threads create file i =
 * blob_host[offsets_host[i] .. offsets_host[i+1]) at paths_host[i] (paths are copied by submit).  wait blocks until
 * every ticket <= `ticket` is on disk; an I/O or CUDA failure is sticky (HIAST_ERR_IO / _CUDA, *errno_out).
 * destroy drains the queue.  Negative return of submit = HIAST_ERR_*.                                            */
HIAST_API int     hiast_writer_create(int n_threads, void** handle_out);
HIAST_API int     hiast_writer_destroy(void* handle);
HIAST_API int64_t hiast_writer_submit(void* handle, const char* const* paths_host, int n_files, const uint8_t* blob_host,
                            size_t blob_host_capacity, const int64_t* offsets_host, size_t bytes_copied,
                            const uint8_t* blob_dev, void* stream);
HIAST_API int     hiast_writer_wait(void* handle, int64_t ticket, int* errno_out);

#ifdef __cplusplus
}
#endif
#endif /* HIAST_B200_H_ */
