/*
 * wsdl_b200.h -- C ABI of the B200-native weak-supervision hot path.
 *
 * The reference (alexncoleman/WeaklySupervisedDL) has no FFI of its own: its hot
 * path is a handful of Python callables that dispatch ATen ops.  Each entry point
 * below replaces the ATen op sequence of one of those callables; the Python
 * mirror in weaklysuperviseddl_b200/ binds them with ctypes (raw device pointers,
 * sizes, the current CUDA stream).  Reference paths are relative to
 * /root/reference/TraditionalModel.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current device unless the name
 *     ends in _host; tensors are dense NCHW;
 *   - the caller owns every buffer, including workspaces; the library allocates
 *     nothing, never synchronises the host, reads no environment variable and keeps
 *     no state between calls except immutable per-process caches (the resolved
 *     cuTensorMapEncodeTiled entry point, "kernel attributes set on device d");
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 ok, <0 argument error (WSDL_E_*), >0 a cudaError_t from
 *     the launch; wsdl_strerror() renders either.  Nothing throws across the ABI.
 */
#ifndef WSDL_B200_H_
#define WSDL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSDL_VERSION 200 /* major*10000 + minor*100 + patch */

#define WSDL_E_NULL (-1)      /* required pointer is NULL */
#define WSDL_E_SHAPE (-2)     /* bad extent (<=0, too many layers/classes, window even or > 2*min(H,W)-1 ...) */
#define WSDL_E_ALIGN (-3)     /* pointer not aligned to its element size */
#define WSDL_E_WORKSPACE (-4) /* workspace too small */
#define WSDL_E_DTYPE (-5)     /* unknown dtype code */
#define WSDL_E_ARG (-6)       /* other invalid argument (alpha_mode, sigma<=0 ...) */

#define WSDL_MAX_LAYERS 8  /* LayerCAM target layers per call */
#define WSDL_MAX_CLASSES 8 /* classes of the pairwise losses */

#define WSDL_F32 0
#define WSDL_BF16 1
#define WSDL_F16 2
#define WSDL_U8 3  /* images as 8-bit pixels (read as value / 255), labels as bytes */
#define WSDL_I64 4 /* labels as int64 (torch.long) */

int wsdl_version(void);
const char* wsdl_strerror(int rc);

/* ---------------------------------------------------------------------------------------
 * Stage 1+2: LayerCAM -> normalise -> upsample -> fuse -> threshold.
 * Replaces LayerCAMGenerator.generate's post-backbone part (LayerCAM.py:52-76; variant
 * AlternatingDirectionCutLoss.py:262-286 with alpha_mode=1) and the threshold idiom
 * `cam[cam < t] = 0; mask = cam > 0` (PsuedoMasks.py:59-62, LayerCAM.py:103-107).
 *
 *   act_host[l], grad_host[l]  device pointers to (B, C[l], h[l], w[l]) of `dtype`
 *   cam_out   nullable, (B, out_h, out_w) f32      -- what generate() returns
 *   mask_out  nullable, (B, out_h, out_w) u8 {0,1} -- (cam >= thresh) && (cam > 0)
 *   near_thresh_count nullable, one u64, INCREMENTED by the number of pixels with
 *             |cam - thresh| < near_band (north_star's "counted and reported" band)
 *   alpha_mode 0: mean of layers -> clamp(0) ** alpha           (LayerCAM.py:74-76)
 *              1: per layer normalise, ** alpha, normalise again (CutLoss.py:271-279)
 * Two kernels are enqueued (channel-sum+min/max+normalise; upsample+fuse+threshold); the
 * full-resolution CAM is never written unless cam_out is given.
 * ------------------------------------------------------------------------------------- */
size_t wsdl_layercam_workspace_bytes(const int* C_host, const int* h_host, const int* w_host, int n_layers, int B,
                                     int dtype);

int wsdl_layercam_fused(const void* const* act_host, const void* const* grad_host, const int* C_host,
                        const int* h_host, const int* w_host, int n_layers, int B, int dtype, int out_h, int out_w,
                        float alpha, int alpha_mode, float thresh, float near_band, float* cam_out,
                        uint8_t* mask_out, unsigned long long* near_thresh_count, void* workspace,
                        size_t workspace_bytes, void* stream);

/* Standalone threshold idiom on an existing CAM (LayerCAM.py:103-107, CutLoss.py:529-536):
 * mask = (cam >= thresh) && (cam > 0), n elements. */
int wsdl_threshold_mask(const float* cam, size_t n, float thresh, float near_band, uint8_t* mask_out,
                        unsigned long long* near_thresh_count, void* stream);

/* keep_largest (PsuedoMasks.py:15-21; duplicate AlternatingDirectionCutLoss.py:206-213):
 * 8-connected components of mask != 0 (skimage.measure.label defaults), keep the largest by area
 * (ties: the component whose first pixel comes first in raster order, i.e. skimage's lowest label).
 * mask/out (B,H,W) u8; out gets {0,1}; an empty mask gives an all-zero out (the reference returns
 * the -- all-zero -- input).  best_area nullable, B u32 (area of the kept component, 0 if none).
 * out may alias mask. */
size_t wsdl_keep_largest_workspace_bytes(int B, int H, int W);

int wsdl_keep_largest(const uint8_t* mask, int B, int H, int W, uint8_t* out, unsigned* best_area, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Stage 3: pairwise regularisers, forward + backward in one launch.
 * Replaces LocalNormalizedCutLoss.forward (AlternatingDirectionCutLoss.py:71-105:
 * inner_softmax=1, sigma_space<=0, divide_by_c=1, per_image_loss=0) and
 * ConstrainToBoundaryLossSingle.forward (AlternatingDirectionBoundaryLoss.py:20-44:
 * inner_softmax=0, sigma_space>0, divide_by_c=0, per_image_loss=1, one image per b).
 *
 *   values (B,C,H,W) f32 : logits (inner_softmax=1) or probabilities
 *   images (B,3,H,W) f32
 *   loss_out  1 float (per_image_loss=0: mean over b,h,w as the reference's .mean())
 *             or B floats (per_image_loss=1: each image's own mean over h,w)
 *   grad_values nullable (B,C,H,W): d(sum of loss_out * grad_out)/d values;
 *             NULL = forward only
 *   grad_out  nullable device pointer to 1 (or B) upstream gradients; NULL = 1.0
 * ------------------------------------------------------------------------------------- */
size_t wsdl_pairwise_workspace_bytes(int B, int H, int W);

int wsdl_pairwise_fwd_bwd(const float* values, const float* images, int B, int C, int H, int W, int window,
                          float sigma_color, float sigma_space, int inner_softmax, int divide_by_c,
                          int per_image_loss, const float* grad_out, float* loss_out, float* grad_values,
                          void* workspace, size_t workspace_bytes, void* stream);

/* The same call for a PREPARED workspace: one that wsdl_pairwise_workspace_init() has initialised once (it zeroes the
 * 512-byte control block and marks the per-CTA result slots behind it "empty" -- all ones; a zero-filled buffer is NOT
 * a prepared workspace) and that has since only been used by complete pairwise calls of ONE problem shape (B, H, W: the
 * regions inside a workspace are laid out per shape) ordered on one stream.
 * Every pairwise kernel leaves the workspace prepared again, so the per-call memsets of wsdl_pairwise_fwd_bwd (extra
 * nodes per call in a CUDA graph) are not needed.  A workspace must not be shared by calls that may run concurrently. */
int wsdl_pairwise_workspace_init(void* workspace, size_t workspace_bytes, void* stream);

int wsdl_pairwise_fwd_bwd_prepared(const float* values, const float* images, int B, int C, int H, int W, int window,
                                   float sigma_color, float sigma_space, int inner_softmax, int divide_by_c,
                                   int per_image_loss, const float* grad_out, float* loss_out, float* grad_values,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* Both regularisers of a two-class batch in ONE pass: LocalNormalizedCutLoss(logits, images)
 * (AlternatingDirectionCutLoss.py:71-105; softmax inside, mean over b,h,w, / C) and, for every image b,
 * ConstrainToBoundaryLossSingle(softmax(logits)[b], images[b]) (AlternatingDirectionBoundaryLoss.py:20-44) -- the
 * composition of the weakly-supervised training step.  Both see the same probability map, so a pixel pair shares its
 * colour distance and p(a) - p(b); only the affinities differ.
 *   logits (B,2,H,W), images (B,3,H,W) f32; window must be 5 and H, W >= 6 (else WSDL_E_SHAPE: use two
 *   wsdl_pairwise_fwd_bwd calls);
 *   loss_cut 1 float, loss_bnd B floats;
 *   grad_logits nullable (B,2,H,W): d(go_cut * loss_cut + sum_b go_bnd[b] * loss_bnd[b]) / d logits, softmax backward
 *   included; grad_out_cut / grad_out_bnd nullable device pointers (1 / B floats), NULL = 1.0;
 *   workspace >= wsdl_pairwise_dual_workspace_bytes(); prepared != 0: initialised with wsdl_pairwise_workspace_init
 *   (self-cleaning, as for wsdl_pairwise_fwd_bwd_prepared), 0: the call resets its result slots itself. */
size_t wsdl_pairwise_dual_workspace_bytes(int B, int H, int W);

int wsdl_pairwise_dual_fwd_bwd(const float* logits, const float* images, int B, int H, int W, int window,
                               float sigma_cut, float sigma_bnd, float sigma_space, const float* grad_out_cut,
                               const float* grad_out_bnd, float* loss_cut, float* loss_bnd, float* grad_logits,
                               void* workspace, size_t workspace_bytes, int prepared, void* stream);

/* The whole loss of a weakly-supervised training step on a two-class batch in ONE launch:
 *     total = lam_ce * CE(logits, labels) + go_cut * cut(logits, images) + sum_b go_bnd[b] * boundary(softmax(logits)[b], images[b])
 * i.e. the criterion of SegmentationModel.py:96-113 (F.cross_entropy, mean over the labels that are not ignore_index)
 * plus the two regularisers above (AlternatingDirectionCutLoss.py:71-105, AlternatingDirectionBoundaryLoss.py:20-44)
 * weighted as AlternatingDirectionCutLoss.py:709 / AlternatingDirectionBoundaryLoss.py:159 weight them.
 *   logits (B,2,H,W) WSDL_F32 | WSDL_BF16 (the network's autocast output; arithmetic is f32 on the values as given);
 *   images (B,3,H,W) WSDL_F32 | WSDL_U8 (the dataset's 8-bit pixels, read as value / 255 exactly like ToTensor);
 *   labels nullable (no CE term), (B,H,W) WSDL_U8 | WSDL_I64, class ids 0 / 1; ignore_index and any other value are skipped;
 *   ce_inv_count nullable device scalar 1 / #(labels != ignore_index) (wsdl_count_valid_labels); NULL = 1 / (B H W);
 *   grad_out_cut / grad_out_bnd nullable device pointers (1 / B floats), NULL = 1.0 -- the lambdas go in here;
 *   loss_ce (nullable without labels), loss_cut: 1 float each, loss_bnd: B floats -- the UNWEIGHTED loss values;
 *   loss_total nullable, 1 float: the weighted total above;
 *   grad_logits nullable (B,2,H,W) WSDL_F32 | WSDL_BF16: d total / d logits;
 *   window must be 5; H, W >= 6; row strides and bases 16-byte aligned (W % 4 == 0 for f32, % 8 for bf16 logits, % 16
 *   for u8 images), else WSDL_E_SHAPE / WSDL_E_ALIGN: compose wsdl_pairwise_fwd_bwd calls instead;
 *   workspace >= wsdl_weak_loss_workspace_bytes(), prepared as for wsdl_pairwise_fwd_bwd_prepared (prepared = 0: the
 *   call zeroes the control block itself).
 * One persistent CTA per SM streams tiles from a queue (csrc/pairwise_stream.cu). */
size_t wsdl_weak_loss_workspace_bytes(int B, int H, int W);

int wsdl_weak_loss_fwd_bwd(const void* logits, int logits_dtype, const void* images, int images_dtype, const void* labels,
                           int labels_dtype, long long ignore_index, int B, int H, int W, int window, float sigma_cut,
                           float sigma_bnd, float sigma_space, float lam_ce, const float* ce_inv_count,
                           const float* grad_out_cut, const float* grad_out_bnd, float* loss_ce, float* loss_cut,
                           float* loss_bnd, float* loss_total, void* grad_logits, int grad_dtype, void* workspace,
                           size_t workspace_bytes, int prepared, void* stream);

/* inv_count[0] = 1 / #(labels[i] != ignore_index), i < n (F.cross_entropy's mean denominator); labels WSDL_U8 | WSDL_I64.
 * scratch: 16 bytes, 8-byte aligned, that the call zeroes and uses; all on `stream`. */
int wsdl_count_valid_labels(const void* labels, int labels_dtype, size_t n, long long ignore_index,
                            unsigned long long* scratch, float* inv_count, void* stream);

/* compute_affinities / compute_affinities_single (AlternatingDirectionCutLoss.py:612-637,
 * AlternatingDirectionBoundaryLoss.py:46-70): images (B,3,H,W) -> out (K,B,H,W) with
 * K = window*window-1, offset order dy outer / dx inner, centre skipped; out[k] is the
 * reference's k-th list entry (B,1,H,W).  sigma_space <= 0 drops the spatial term. */
int wsdl_affinities(const float* images, int B, int H, int W, int window, float sigma_color, float sigma_space,
                    float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Either side of the path.
 * wsdl_labels_from_masks: masks (B,H,W) u8 {0,1} -> training labels (B,S,S) i64 {0,1}; what the reference gets from
 * save_image (PsuedoMasks.py:67-69) -> PNG -> convert('L') + NEAREST resize + int64 (SegmentationDataset.py:21,26,35)
 * -> clamp(max=1) (SegmentationModel.py:100), without the PNG round trip.
 * wsdl_iou_acc_counts: counts[b] += {|pred>0 & true>0|, |pred>0 | true>0|, |pred == true|} over the n elements of
 * image b (compute_iou_and_acc, ExtraUtilities.py:4-21); elem: 0 u8, 1 i32, 2 i64, 3 f32; counts (B,3) u64, zeroed by
 * the caller.
 * ------------------------------------------------------------------------------------- */
int wsdl_labels_from_masks(const uint8_t* mask, int B, int H, int W, int out_size, long long* labels, void* stream);

int wsdl_iou_acc_counts(const void* pred, const void* truth, int B, size_t n, int elem, unsigned long long* counts,
                        void* stream);

/* refine_pseudo_mask (AlternatingDirectionCutLoss.py:709-767) for a batch of independent images, the element-wise half
 * of a step: X (B,2,H,W) free variable, S = softmax(model(image)) (B,2,H,W).  One call FINISHES step `step` and BEGINS
 * the next one in a single pass:
 *   with g_cut != NULL: d/dXn of  KL(S || Xn) + lam * kl_in[b] / (loss_cut[b] + 1e-6) * cut_b  (g_cut = d cut_b / d Xn for an
 *     upstream gradient of 1, loss_cut: the per-image cut losses -- both from one wsdl_pairwise_fwd_bwd call on the Xn this
 *     function wrote last time, inner_softmax = 1, divide_by_c = 1, per_image_loss = 1), through the softmax, Adam update
 *     (torch defaults' arithmetic; `step` is Adam's 1-based step count) of X, exp_avg, exp_avg_sq in place;
 *   always: Xn_out = softmax(X), kl_out[b] = sum S (log S - log(Xn + 1e-8)) of image b (:737-740), and, when mask_out is
 *     given, mask_out (B,H,W) = (Xn[:,1] > threshold) as 0 / 1 floats (:759-760).
 * The first call of a refinement passes g_cut = NULL (nothing to finish yet).  A step is this launch + the cut launch:
 * the dynamic weight (:748, two host reads per step in the reference) never leaves the device.
 * workspace >= wsdl_refine_workspace_bytes(), prepared as for the pairwise calls (0: the call zeroes its tickets). */
size_t wsdl_refine_workspace_bytes(int B, int H, int W);

int wsdl_refine_step(float* X, const float* S, const float* g_cut, const float* loss_cut, const float* kl_in, float* exp_avg,
                     float* exp_avg_sq, float* Xn_out, float* kl_out, float* mask_out, int B, int H, int W, float lam,
                     float lr, float beta1, float beta2, float eps, int step, float threshold, void* workspace,
                     size_t workspace_bytes, int prepared, void* stream);

/* Scale of a saved gradient by a device scalar (autograd backward of the fused launch, whose
 * gradient was computed for an upstream gradient of 1): dst[i] = src[i] * scale[per ? i / per : 0].
 * dst may alias src. */
int wsdl_scale(const float* src, float* dst, size_t n, const float* scale, size_t per, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WSDL_B200_H_ */
