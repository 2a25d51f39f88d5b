// Parameters and helpers shared by the pairwise-regulariser kernels (pairwise.cu, pairwise_sym.cu).
#pragma once
#include "common.cuh"

namespace wsdl {

constexpr float LOG2E = 1.4426950408889634f;

struct PwParams {
  const float* values;
  const float* images;
  const float* grad_out;  // nullable
  float* loss_out;
  float* grad_values;     // nullable
  float* partial;         // block partial sums (layout is the kernel's own)
  unsigned* ticket;
  unsigned* sym_slots;    // pair-symmetric kernels: one result slot per CTA, 0xffffffff when empty (a region of its own)
  int B, C, H, W, pad;
  int tiles_x, tiles_y;
  int inner_softmax, per_image;
  float kc;        // -log2(e) / (2 sigma_color^2)
  float ks_unit;   // -log2(e) / (2 sigma_space^2), 0 when there is no spatial term
  float inv_2ss;   // 1 / (2 sigma_space^2), 0 when there is no spatial term
  double kappa;    // 1 / (K * N * (C or 1)), N = B*H*W or H*W
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// Pair-symmetric window-5 kernel (pairwise_sym.cu).  ps_launch returns 1 when the shape is not its own
// (the caller then takes the generic kernels), 0 on success, another value on a CUDA error.
int ps_launch(const PwParams& P, cudaStream_t s);
size_t ps_workspace_floats(int B, int H, int W);

// Cut loss on the logits + boundary loss on softmax(logits) of a two-class batch in one pass (pairwise_sym.cu).
// P describes the cut loss; returns 1 when the shape is not supported.
int ps_launch_dual(const PwParams& P, float sigma_cut, float sigma_bnd, float sigma_space, const float* grad_out_bnd,
                   float* loss_bnd, unsigned long long* slots, cudaStream_t s);

// Persistent tile-streaming kernel for the whole weak-supervision loss of a two-class batch (pairwise_stream.cu):
// cut + boundary (+ cross-entropy) forward and d/dlogits in one launch, f32 / bf16 logits, f32 / u8 images.
struct SgLaunch {
  const void* logits;   // (B,2,H,W) logit_dtype
  const void* images;   // (B,3,H,W) image_dtype (u8: 0..255, read as value / 255)
  const void* labels;   // nullable (B,H,W) label_dtype
  void* grad;           // nullable (B,2,H,W) grad_dtype
  const float* go_cut;  // nullable, 1
  const float* go_bnd;  // nullable, B
  const float* ce_inv_n_dev;  // nullable, 1
  float* loss_cut;      // 1
  float* loss_bnd;      // B
  float* loss_ce;       // 1 (nullable)
  float* loss_total;    // 1 (nullable)
  float* partial;       // sg_workspace_floats()
  unsigned* ctrl;       // 2 words, zero before the call, zero after it
  long long ignore_index;
  int B, H, W;
  int logit_dtype, image_dtype, label_dtype, grad_dtype;
  float sigma_cut, sigma_bnd, sigma_space, lam_ce;
};
int sg_launch(const SgLaunch& L, cudaStream_t s);  // 1: not its shape, 0: ok, else a CUDA error
size_t sg_workspace_floats(int B, int H, int W);

}  // namespace wsdl
