// Alternating-direction refinement of pseudo-masks (refine_pseudo_mask, reference TraditionalModel/
// AlternatingDirectionCutLoss.py:709-767), the element-wise half of a step, for a batch of independent images.
//
// One step of the reference, per image (X: (2,H,W) free variable, S = softmax(model(image)) fixed):
//     Xn   = softmax(X)                                                        :737
//     kl   = sum S (log S - log(Xn + 1e-8))         (kl_div, 'batchmean', batch of one)   :740
//     cut  = LocalNormalizedCutLoss(Xn, image)      (its own softmax inside)   :745 -> :78
//     lam' = lam * kl / (cut + 1e-6)                (two .item() reads, python doubles)    :748
//     (kl + lam' * cut).backward();  Adam step on X :750-756
// The cut loss and its gradient are one launch of the pairwise kernel (wsdl_pairwise_fwd_bwd_prepared, per-image
// losses).  Everything else is this kernel, which FINISHES step t and BEGINS step t + 1 in one pass over X:
//     dKL/dXn = -S / (Xn + 1e-8),  g = dKL/dXn + lam' dcut/dXn,  dX = Xn (g - sum_c g_c Xn_c)   (softmax backward)
//     Adam (torch defaults) on X, m, v
//     Xn' = softmax(X'),  kl' = per-image sum (block partials, last block of an image adds them in a fixed order)
// so a step is two launches and nothing ever goes to the host: lam' is formed on the device from the two per-image
// scalars, in double and rounded to float exactly where python hands its double to the float tensor.
#include "common.cuh"

namespace wsdl {

constexpr int RF_THREADS = 256;
constexpr int RF_PIX = 4;  // pixels per thread (one float4 per plane)

struct RfParams {
  float* X;             // (B,2,H,W), updated in place
  const float* S;       // (B,2,H,W)
  const float* g_cut;   // (B,2,H,W) d cut_b / d Xn for an upstream gradient of 1; NULL: no update (first call)
  const float* loss_cut;  // B
  const float* kl_in;     // B: kl of the Xn that g_cut was computed on
  float* m;             // Adam exp_avg
  float* v;             // Adam exp_avg_sq
  float* Xn;            // (B,2,H,W) out: softmax of the updated X
  float* kl_out;        // B out
  float* mask_out;      // nullable (B,H,W): (Xn[:,1] > threshold) as 0 / 1 floats
  float* partial;       // [B][blocks_per_image]
  unsigned* ticket;     // [B], zero before and after
  int B, n_pix, blocks_per_image;
  float lam, threshold;
  float beta2, eps;
  float w1, w2;         // 1 - beta1, 1 - beta2: python doubles rounded to float
  float step_size;      // lr / (1 - beta1^t)
  float bc2_sqrt;       // sqrt(1 - beta2^t)
};

__global__ void __launch_bounds__(RF_THREADS) refine_step_kernel(const __grid_constant__ RfParams P) {
  __shared__ float s_red[RF_THREADS / 32];
  __shared__ int s_last;
  const int b = blockIdx.y;
  const size_t plane = (size_t)P.n_pix, base = (size_t)b * 2 * plane;
  const bool update = P.g_cut != nullptr;
  pdl_wait();  // launched with programmatic stream serialization: scheduled under the tail of the cut launch before it
  pdl_launch_dependents();
  float lamp = 0.f;
  if (update) {  // lam * (kl / (cut + 1e-6)) in double, handed to the float tensor as a float (python scalar semantics)
    const double kl = (double)__ldg(P.kl_in + b), cut = (double)__ldg(P.loss_cut + b);
    lamp = (float)((double)P.lam * (kl / (cut + 1e-6)));
  }
  float kl_sum = 0.f;
  const int i0 = (blockIdx.x * RF_THREADS + threadIdx.x) * RF_PIX;
  if (i0 < P.n_pix) {
    const bool vec = (i0 + RF_PIX <= P.n_pix) && ((P.n_pix & 3) == 0);
    float x[2][RF_PIX], s[2][RF_PIX];
    const int n = min(RF_PIX, P.n_pix - i0);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (vec) {
        const float4 a = *reinterpret_cast<const float4*>(P.X + base + c * plane + i0);
        const float4 t = __ldg(reinterpret_cast<const float4*>(P.S + base + c * plane + i0));
        x[c][0] = a.x, x[c][1] = a.y, x[c][2] = a.z, x[c][3] = a.w;
        s[c][0] = t.x, s[c][1] = t.y, s[c][2] = t.z, s[c][3] = t.w;
      } else {
#pragma unroll
        for (int j = 0; j < RF_PIX; ++j) {
          x[c][j] = j < n ? P.X[base + c * plane + i0 + j] : 0.f;
          s[c][j] = j < n ? __ldg(P.S + base + c * plane + i0 + j) : 0.f;
        }
      }
    }
    if (update) {
      float g[2][RF_PIX], m[2][RF_PIX], v[2][RF_PIX];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (vec) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(P.g_cut + base + c * plane + i0));
          const float4 mm = *reinterpret_cast<const float4*>(P.m + base + c * plane + i0);
          const float4 vv = *reinterpret_cast<const float4*>(P.v + base + c * plane + i0);
          g[c][0] = a.x, g[c][1] = a.y, g[c][2] = a.z, g[c][3] = a.w;
          m[c][0] = mm.x, m[c][1] = mm.y, m[c][2] = mm.z, m[c][3] = mm.w;
          v[c][0] = vv.x, v[c][1] = vv.y, v[c][2] = vv.z, v[c][3] = vv.w;
        } else {
#pragma unroll
          for (int j = 0; j < RF_PIX; ++j) {
            g[c][j] = j < n ? __ldg(P.g_cut + base + c * plane + i0 + j) : 0.f;
            m[c][j] = j < n ? P.m[base + c * plane + i0 + j] : 0.f;
            v[c][j] = j < n ? P.v[base + c * plane + i0 + j] : 0.f;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < RF_PIX; ++j) {
        // Xn of the OLD X (what the cut launch saw): recomputed, two classes
        const float mx = fmaxf(x[0][j], x[1][j]);
        const float e0 = expf(x[0][j] - mx), e1 = expf(x[1][j] - mx);
        const float p0 = __fdiv_rn(e0, e0 + e1), p1 = __fdiv_rn(e1, e0 + e1);
        const float g0 = __fdiv_rn(-s[0][j], p0 + 1e-8f) + lamp * g[0][j];
        const float g1 = __fdiv_rn(-s[1][j], p1 + 1e-8f) + lamp * g[1][j];
        const float dot = g0 * p0 + g1 * p1;
        const float d[2] = {p0 * (g0 - dot), p1 * (g1 - dot)};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          m[c][j] = m[c][j] + (d[c] - m[c][j]) * P.w1;                   // exp_avg.lerp_(grad, 1 - beta1)
          v[c][j] = fmaf(d[c] * d[c], P.w2, v[c][j] * P.beta2);           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
          const float denom = __fdiv_rn(sqrtf(v[c][j]), P.bc2_sqrt) + P.eps;
          x[c][j] = x[c][j] - P.step_size * __fdiv_rn(m[c][j], denom);               // addcdiv_(m, denom, -step_size)
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (vec) {
          *reinterpret_cast<float4*>(P.X + base + c * plane + i0) = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
          *reinterpret_cast<float4*>(P.m + base + c * plane + i0) = make_float4(m[c][0], m[c][1], m[c][2], m[c][3]);
          *reinterpret_cast<float4*>(P.v + base + c * plane + i0) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        } else {
#pragma unroll
          for (int j = 0; j < RF_PIX; ++j)
            if (j < n) {
              P.X[base + c * plane + i0 + j] = x[c][j];
              P.m[base + c * plane + i0 + j] = m[c][j];
              P.v[base + c * plane + i0 + j] = v[c][j];
            }
        }
      }
    }
    // begin the next step: Xn of the new X and its KL terms
    float q[2][RF_PIX];
#pragma unroll
    for (int j = 0; j < RF_PIX; ++j) {
      const float mx = fmaxf(x[0][j], x[1][j]);
      const float e0 = expf(x[0][j] - mx), e1 = expf(x[1][j] - mx);
      q[0][j] = __fdiv_rn(e0, e0 + e1), q[1][j] = __fdiv_rn(e1, e0 + e1);
      if (j < n) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float t = s[c][j];
          const float xlogy = t == 0.f ? 0.f : t * logf(t);              // torch.xlogy(S, S)
          kl_sum += xlogy - t * logf(q[c][j] + 1e-8f);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (vec) {
        *reinterpret_cast<float4*>(P.Xn + base + c * plane + i0) = make_float4(q[c][0], q[c][1], q[c][2], q[c][3]);
      } else {
#pragma unroll
        for (int j = 0; j < RF_PIX; ++j)
          if (j < n) P.Xn[base + c * plane + i0 + j] = q[c][j];
      }
    }
    if (P.mask_out) {
#pragma unroll
      for (int j = 0; j < RF_PIX; ++j)
        if (j < n) P.mask_out[(size_t)b * plane + i0 + j] = q[1][j] > P.threshold ? 1.f : 0.f;
    }
  }
  // per-image KL: block partial, the last block of the image adds the partials in block order
  kl_sum = warp_sum(kl_sum);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = kl_sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < RF_THREADS / 32; ++i) t += s_red[i];
    __stcg(P.partial + (size_t)b * P.blocks_per_image + blockIdx.x, t);
    __threadfence();
    s_last = atomicAdd(P.ticket + b, 1u) == (unsigned)P.blocks_per_image - 1u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < P.blocks_per_image; i += RF_THREADS)
    acc += (double)ld_cg_f32(P.partial + (size_t)b * P.blocks_per_image + i);
  __shared__ double s_d[RF_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_d[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < RF_THREADS / 32; ++i) t += s_d[i];
    P.kl_out[b] = (float)t;
    P.ticket[b] = 0u;  // ready for the next launch
  }
}

}  // namespace wsdl

using namespace wsdl;

static int rf_blocks(int n_pix) { return (n_pix + RF_THREADS * RF_PIX - 1) / (RF_THREADS * RF_PIX); }

extern "C" size_t wsdl_refine_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t tickets = ((size_t)B * 4 + 255) / 256 * 256;
  return 256 + tickets + (size_t)B * rf_blocks(H * W) * 4;
}

extern "C" int wsdl_refine_step(float* X, const float* S, const float* g_cut, const float* loss_cut, const float* kl_in,
                                float* exp_avg, float* exp_avg_sq, float* Xn_out, float* kl_out, float* mask_out, int B,
                                int H, int W, float lam, float lr, float beta1, float beta2, float eps, int step,
                                float threshold, void* workspace, size_t workspace_bytes, int prepared, void* stream) {
  if (!X || !S || !Xn_out || !kl_out || !workspace) return WSDL_E_NULL;
  if (g_cut && (!loss_cut || !kl_in || !exp_avg || !exp_avg_sq)) return WSDL_E_NULL;
  if (B < 1 || B > 65535 || H < 1 || W < 1 || (long long)H * W > (1LL << 30)) return WSDL_E_SHAPE;
  if (g_cut && step < 1) return WSDL_E_ARG;
  if (workspace_bytes < wsdl_refine_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  const bool vec = ((H * W) & 3) == 0;
  const uintptr_t need = vec ? 15 : 3;
  const void* ptrs[] = {X, S, g_cut, exp_avg, exp_avg_sq, Xn_out};
  for (const void* p : ptrs)
    if (p && ((uintptr_t)p & need)) return WSDL_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  RfParams P;
  P.X = X, P.S = S, P.g_cut = g_cut, P.loss_cut = loss_cut, P.kl_in = kl_in, P.m = exp_avg, P.v = exp_avg_sq;
  P.Xn = Xn_out, P.kl_out = kl_out, P.mask_out = mask_out;
  const uintptr_t w0 = ((uintptr_t)workspace + 255) / 256 * 256;
  P.ticket = reinterpret_cast<unsigned*>(w0);
  const size_t tickets = ((size_t)B * 4 + 255) / 256 * 256;
  P.partial = reinterpret_cast<float*>(w0 + tickets);
  P.B = B, P.n_pix = H * W, P.blocks_per_image = rf_blocks(H * W);
  P.lam = lam, P.threshold = threshold, P.beta2 = beta2, P.eps = eps;
  P.w1 = (float)(1.0 - (double)beta1), P.w2 = (float)(1.0 - (double)beta2);
  const int t = step < 1 ? 1 : step;
  // python doubles, handed to the tensor ops as floats (torch.optim.Adam, single-tensor path)
  P.step_size = (float)((double)lr / (1.0 - pow((double)beta1, (double)t)));
  P.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)t));
  if (!prepared) {
    cudaError_t e = cudaMemsetAsync(P.ticket, 0, tickets, s);
    if (e != cudaSuccess) return (int)e;
  }
  {
    const cudaError_t e = launch_pdl(refine_step_kernel, dim3(P.blocks_per_image, B), dim3(RF_THREADS), 0, s, P);
    if (e != cudaSuccess) return (int)e;
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}
