// Pairwise regularisers (cut loss / boundary loss), forward + backward in ONE launch, sm_100a.
//
// Replaces LocalNormalizedCutLoss.forward (reference TraditionalModel/
// AlternatingDirectionCutLoss.py:71-105) and ConstrainToBoundaryLossSingle.forward
// (AlternatingDirectionBoundaryLoss.py:20-70) plus their autograd backward, i.e. ~3 900 ATen
// launches per call, with a single stencil kernel that reads values+image once and writes the
// gradient once (28 B/pixel at C=2).
//
// Math.  With r() the reflect padding, k(z,y)=exp(-|I(z)-I(y)|^2/(2 sc^2)) and
// M(z,y) = sum over window offsets d of [r(z+d)==y] * s(d)   (s = spatial Gaussian or 1)
// the reference's loss is          L = kappa * sum_z sum_y M(z,y) k(z,y) |p(z)-p(y)|^2
// and its gradient (gather form)   dL/dp_c(z) = 2 kappa * sum_y [M(z,y)+M(y,z)] k(z,y) (p_c(z)-p_c(y)),
// y ranging over the in-image window around z.  M is separable, M = my(zy,ey)*mx(zx,ex), and equals
// s(e) whenever z is >= pad pixels from every border, so CTAs whose tile is >= 2*pad from the border
// take a table-free path.  With an inner softmax, dL/dlogit_c = p_c (g_c - sum_j p_j g_j).
#include "common.cuh"

namespace wsdl {

constexpr int PW_TW = 32;       // tile width  (one lane per column)
constexpr int PW_TH = 32;       // tile height
constexpr int PW_THREADS = 256; // 8 warps, warp w owns rows w, w+8, w+16, w+24
constexpr int PW_MAXPAD = 3;    // window <= 7
constexpr float LOG2E = 1.4426950408889634f;

struct PwParams {
  const float* values;
  const float* images;
  const float* grad_out;  // nullable
  float* loss_out;
  float* grad_values;     // nullable
  float* partial;         // (B * tiles) block partial sums
  unsigned* ticket;
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

// M_axis(a, e; n) = sum_{d=-pad..pad} [reflect(a+d) == a+e] * s1(d); 0 when a or a+e is outside.
__device__ float axis_multiplicity(int a, int e, int n, int pad, float inv_2ss) {
  if (a < 0 || a >= n || a + e < 0 || a + e >= n) return 0.f;
  float m = 0.f;
  for (int d = -pad; d <= pad; ++d)
    if (reflect_idx(a + d, n) == a + e) m += (inv_2ss != 0.f) ? expf(-(float)(d * d) * inv_2ss) : 1.f;
  return m;
}

template <int CT, int PADT>  // CT = 0 / PADT = 0: runtime C (<= 8) / runtime pad (<= 3)
__global__ void __launch_bounds__(PW_THREADS) pairwise_fwd_bwd_kernel(const __grid_constant__ PwParams P) {
  constexpr int CMAX = CT ? CT : WSDL_MAX_CLASSES;
  constexpr int PMAX = PADT ? PADT : PW_MAXPAD;
  constexpr int SW = PW_TW + 2 * PMAX;  // smem row stride
  constexpr int SH = PW_TH + 2 * PMAX;
  const int C = CT ? CT : P.C;
  const int pad = PADT ? PADT : P.pad;
  const int win = 2 * pad + 1;

  extern __shared__ float smem[];
  float* s_img = smem;                      // [3][SH][SW]
  float* s_p = s_img + 3 * SH * SW;         // [C][SH][SW]
  float* s_tab = s_p + CMAX * SH * SW;      // myf[TH][7] myb[TH][7] mxf[TW][7] mxb[TW][7]
  __shared__ float s_red[PW_THREADS / 32];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y, b = blockIdx.z;
  const int x0 = tile_x * PW_TW, y0 = tile_y * PW_TH;
  const int H = P.H, W = P.W;
  const size_t plane = (size_t)H * W;
  const float* img = P.images + (size_t)b * 3 * plane;
  const float* val = P.values + (size_t)b * C * plane;

  // ---- stage tile + halo (zeros outside the image; softmax applied on the way in) ----
  const int lw = PW_TW + 2 * pad, lh = PW_TH + 2 * pad;
  for (int i = tid; i < lw * lh; i += PW_THREADS) {
    const int ty = i / lw, tx = i - ty * lw;
    const int gy = y0 - pad + ty, gx = x0 - pad + tx;
    const bool in = (gy >= 0 && gy < H && gx >= 0 && gx < W);
    const size_t o = in ? (size_t)gy * W + gx : 0;
    const int so = ty * SW + tx;
#pragma unroll
    for (int c = 0; c < 3; ++c) s_img[c * SH * SW + so] = in ? __ldg(img + c * plane + o) : 0.f;
    float v[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) v[c] = (c < C && in) ? __ldg(val + c * plane + o) : 0.f;
    if (P.inner_softmax) {
      float m = v[0];
#pragma unroll
      for (int c = 1; c < CMAX; ++c)
        if (c < C) m = fmaxf(m, v[c]);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < C) {
          v[c] = ex2_approx((v[c] - m) * LOG2E);
          s += v[c];
        }
      const float inv = __fdiv_rn(1.f, s);
#pragma unroll
      for (int c = 0; c < CMAX; ++c) v[c] = in ? v[c] * inv : 0.f;
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
      if (c < C) s_p[c * SH * SW + so] = v[c];
  }

  const bool border = (x0 < 2 * pad) || (y0 < 2 * pad) || (x0 + PW_TW > W - 2 * pad) || (y0 + PW_TH > H - 2 * pad);
  constexpr int TS = 2 * PW_MAXPAD + 1;
  float* s_myf = s_tab;
  float* s_myb = s_myf + PW_TH * TS;
  float* s_mxf = s_myb + PW_TH * TS;
  float* s_mxb = s_mxf + PW_TW * TS;
  if (border) {
    for (int i = tid; i < PW_TH * win; i += PW_THREADS) {
      const int rr = i / win, e = i - rr * win - pad;
      s_myf[rr * TS + e + pad] = axis_multiplicity(y0 + rr, e, H, pad, P.inv_2ss);
      s_myb[rr * TS + e + pad] = axis_multiplicity(y0 + rr + e, -e, H, pad, P.inv_2ss);
    }
    for (int i = tid; i < PW_TW * win; i += PW_THREADS) {
      const int cc = i / win, e = i - cc * win - pad;
      s_mxf[cc * TS + e + pad] = axis_multiplicity(x0 + cc, e, W, pad, P.inv_2ss);
      s_mxb[cc * TS + e + pad] = axis_multiplicity(x0 + cc + e, -e, W, pad, P.inv_2ss);
    }
  }
  __syncthreads();

  const float scale_g = (float)(2.0 * P.kappa) * (P.grad_out ? __ldg(P.grad_out + (P.per_image ? b : 0)) : 1.f);
  float loss_acc = 0.f;
  const int gx = x0 + lane;

  for (int rr = warp; rr < PW_TH; rr += PW_THREADS / 32) {
    const int gy = y0 + rr;
    if (gy >= H) break;  // warp-uniform
    const int so = (rr + pad) * SW + (lane + pad);
    const float i0 = s_img[so], i1 = s_img[SH * SW + so], i2 = s_img[2 * SH * SW + so];
    float pc[CMAX], g[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      pc[c] = (c < C) ? s_p[c * SH * SW + so] : 0.f;
      g[c] = 0.f;
    }
    float lpix = 0.f;
#pragma unroll
    for (int ey = -PMAX; ey <= PMAX; ++ey) {
      if (ey < -pad || ey > pad) continue;
      float myf = 1.f, myb = 1.f;
      if (border) {
        myf = s_myf[rr * TS + ey + pad];
        myb = s_myb[rr * TS + ey + pad];
      }
#pragma unroll
      for (int ex = -PMAX; ex <= PMAX; ++ex) {
        if (ex < -pad || ex > pad || (ex == 0 && ey == 0)) continue;
        const int sn = so + ey * SW + ex;
        const float d0 = i0 - s_img[sn], d1 = i1 - s_img[SH * SW + sn], d2 = i2 - s_img[2 * SH * SW + sn];
        const float dist = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
        float cf, cg, k;
        if (border) {
          k = ex2_approx(dist * P.kc);
          cf = myf * s_mxf[lane * TS + ex + pad];
          cg = fmaf(myb, s_mxb[lane * TS + ex + pad], cf);
        } else {
          k = ex2_approx(fmaf(dist, P.kc, (float)(ex * ex + ey * ey) * P.ks_unit));
          cf = 1.f;
          cg = 2.f;
        }
        float sq = 0.f;
        const float kg = k * cg;
#pragma unroll
        for (int c = 0; c < CMAX; ++c)
          if (c < C) {
            const float dp = pc[c] - s_p[c * SH * SW + sn];
            sq = fmaf(dp, dp, sq);
            g[c] = fmaf(kg, dp, g[c]);
          }
        lpix = fmaf(k * cf, sq, lpix);
      }
    }
    if (gx < W) {
      loss_acc += lpix;
      if (P.grad_values) {
        float* go = P.grad_values + (size_t)b * C * plane + (size_t)gy * W + gx;
        if (P.inner_softmax) {
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) dot = fmaf(pc[c], g[c], dot);
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) go[c * plane] = scale_g * pc[c] * (g[c] - dot);
        } else {
#pragma unroll
          for (int c = 0; c < CMAX; ++c)
            if (c < C) go[c * plane] = scale_g * g[c];
        }
      }
    }
  }

  // ---- loss: block partial -> workspace; the last CTA adds everything up in a fixed order ----
  loss_acc = warp_sum(loss_acc);
  if (lane == 0) s_red[warp] = loss_acc;
  __syncthreads();
  const int tiles = P.tiles_x * P.tiles_y;
  if (tid == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < PW_THREADS / 32; ++i) t += s_red[i];
    __stcg(P.partial + (size_t)b * tiles + tile_y * P.tiles_x + tile_x, t);
    __threadfence();
    const unsigned n = atomicAdd(P.ticket, 1u);
    s_last = (n == (unsigned)(tiles * P.B) - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (P.per_image) {
    for (int bb = warp; bb < P.B; bb += PW_THREADS / 32) {
      double acc = 0.0;
      for (int i = lane; i < tiles; i += 32) acc += (double)ld_cg_f32(P.partial + (size_t)bb * tiles + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) P.loss_out[bb] = (float)(acc * P.kappa);
    }
  } else {
    __shared__ double s_d[PW_THREADS / 32];
    double acc = 0.0;
    const int n = tiles * P.B;
    for (int i = tid; i < n; i += PW_THREADS) acc += (double)ld_cg_f32(P.partial + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_d[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < PW_THREADS / 32; ++i) t += s_d[i];
      P.loss_out[0] = (float)(t * P.kappa);
    }
  }
}

template <int CT, int PADT>
static size_t pw_smem_bytes() {
  constexpr int CMAX = CT ? CT : WSDL_MAX_CLASSES;
  constexpr int PMAX = PADT ? PADT : PW_MAXPAD;
  constexpr int SW = PW_TW + 2 * PMAX, SH = PW_TH + 2 * PMAX;
  return sizeof(float) * ((size_t)(3 + CMAX) * SH * SW + 2 * (PW_TH + PW_TW) * (2 * PW_MAXPAD + 1));
}

template <int CT, int PADT>
static int pw_launch(const PwParams& P, cudaStream_t s) {
  const size_t smem = pw_smem_bytes<CT, PADT>();
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_fwd_bwd_kernel<CT, PADT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(P.tiles_x, P.tiles_y, P.B);
  pairwise_fwd_bwd_kernel<CT, PADT><<<grid, PW_THREADS, smem, s>>>(P);
  WSDL_LAUNCH_CHECK();
  return 0;
}

// ---- compute_affinities: (B,3,H,W) -> (K,B,H,W), reference op order, accurate expf ----
// `two_sc2` is float32(2*sigma_color**2) and the spatial term float32(dist / (2*sigma_space**2)) evaluated in
// double, which is how eager torch folds the python scalars of the reference expression.
__global__ void affinities_kernel(const float* __restrict__ img, int B, int H, int W, int pad, float two_sc2,
                                  double two_ss2, int spatial, float* __restrict__ out) {
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  const int win = 2 * pad + 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / plane);
    const size_t r = i - (size_t)b * plane;
    const int y = (int)(r / W), x = (int)(r - (size_t)y * W);
    const float* ib = img + (size_t)b * 3 * plane;
    const float c0 = ib[r], c1 = ib[plane + r], c2 = ib[2 * plane + r];
    int k = 0;
    for (int dy = -pad; dy <= pad; ++dy)
      for (int dx = -pad; dx <= pad; ++dx) {
        if (dx == 0 && dy == 0) continue;
        const size_t o = (size_t)reflect_idx(y + dy, H) * W + reflect_idx(x + dx, W);
        const float d0 = c0 - ib[o], d1 = c1 - ib[plane + o], d2 = c2 - ib[2 * plane + o];
        const float diff = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
        float arg = __fdiv_rn(-diff, two_sc2);  // -diff / (2 sigma_color^2)
        if (spatial) arg = __fsub_rn(arg, (float)((double)(dx * dx + dy * dy) / two_ss2));
        out[(size_t)k * total + i] = expf(arg);
        ++k;
      }
    (void)win;
  }
}

__global__ void scale_kernel(const float* src, float* dst, size_t n, const float* __restrict__ scale, size_t per) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i] * __ldg(scale + (per ? i / per : 0));
}

}  // namespace wsdl

using namespace wsdl;

static size_t pw_align(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t wsdl_pairwise_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t tiles = (size_t)((W + PW_TW - 1) / PW_TW) * ((H + PW_TH - 1) / PW_TH);
  return 512 + pw_align((size_t)B * tiles * sizeof(float));
}

extern "C" int wsdl_pairwise_fwd_bwd(const float* values, const float* images, int B, int C, int H, int W,
                                     int window, float sigma_color, float sigma_space, int inner_softmax,
                                     int divide_by_c, int per_image_loss, const float* grad_out, float* loss_out,
                                     float* grad_values, void* workspace, size_t workspace_bytes, void* stream) {
  if (!values || !images || !loss_out || !workspace) return WSDL_E_NULL;
  if (B < 1 || H < 1 || W < 1 || C < 1 || C > WSDL_MAX_CLASSES || B > 65535) return WSDL_E_SHAPE;
  if (window < 1 || (window & 1) == 0 || window / 2 > PW_MAXPAD) return WSDL_E_SHAPE;
  const int pad = window / 2;
  if (pad >= H || pad >= W) return WSDL_E_SHAPE;  // reflect padding needs pad < size (as F.pad does)
  if (!(sigma_color > 0.f)) return WSDL_E_ARG;
  if (((uintptr_t)values % 4) || ((uintptr_t)images % 4) || ((uintptr_t)loss_out % 4) ||
      (grad_values && ((uintptr_t)grad_values % 4)))
    return WSDL_E_ALIGN;
  if (workspace_bytes < wsdl_pairwise_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  if (window == 1) {  // no neighbours: the reference divides 0.0 by count 0 -> ZeroDivisionError
    return WSDL_E_SHAPE;
  }
  PwParams P;
  P.values = values;
  P.images = images;
  P.grad_out = grad_out;
  P.loss_out = loss_out;
  P.grad_values = grad_values;
  uintptr_t ws = ((uintptr_t)workspace + 255) / 256 * 256;
  P.ticket = reinterpret_cast<unsigned*>(ws);
  P.partial = reinterpret_cast<float*>(ws + 256);
  P.B = B, P.C = C, P.H = H, P.W = W, P.pad = pad;
  P.tiles_x = (W + PW_TW - 1) / PW_TW;
  P.tiles_y = (H + PW_TH - 1) / PW_TH;
  if (P.tiles_y > 65535) return WSDL_E_SHAPE;
  P.inner_softmax = inner_softmax ? 1 : 0;
  P.per_image = per_image_loss ? 1 : 0;
  P.kc = -LOG2E / (2.f * sigma_color * sigma_color);
  const bool spatial = sigma_space > 0.f;
  P.inv_2ss = spatial ? 1.f / (2.f * sigma_space * sigma_space) : 0.f;
  P.ks_unit = spatial ? -LOG2E * P.inv_2ss : 0.f;
  const double K = (double)window * window - 1.0;
  const double N = (per_image_loss ? 1.0 : (double)B) * (double)H * (double)W;
  P.kappa = 1.0 / (K * N * (divide_by_c ? (double)C : 1.0));
  cudaError_t e = cudaMemsetAsync(P.ticket, 0, 4, s);
  if (e != cudaSuccess) return (int)e;
  if (pad == 2) {
    switch (C) {
      case 1: return pw_launch<1, 2>(P, s);
      case 2: return pw_launch<2, 2>(P, s);
      case 3: return pw_launch<3, 2>(P, s);
      case 4: return pw_launch<4, 2>(P, s);
      default: break;
    }
  }
  return pw_launch<0, 0>(P, s);
}

extern "C" int wsdl_affinities(const float* images, int B, int H, int W, int window, float sigma_color,
                               float sigma_space, float* out, void* stream) {
  if (!images || !out) return WSDL_E_NULL;
  if (B < 1 || H < 1 || W < 1) return WSDL_E_SHAPE;
  if (window < 3 || (window & 1) == 0) return WSDL_E_SHAPE;
  const int pad = window / 2;
  if (pad >= H || pad >= W) return WSDL_E_SHAPE;
  if (!(sigma_color > 0.f)) return WSDL_E_ARG;
  const size_t total = (size_t)B * H * W;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)WSDL_NUM_SMS * 8) blocks = (size_t)WSDL_NUM_SMS * 8;
  const bool spatial = sigma_space > 0.f;
  affinities_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      images, B, H, W, pad, (float)(2.0 * (double)sigma_color * (double)sigma_color),
      spatial ? 2.0 * (double)sigma_space * (double)sigma_space : 1.0, spatial ? 1 : 0, out);
  WSDL_LAUNCH_CHECK();
  return 0;
}

extern "C" int wsdl_scale(const float* src, float* dst, size_t n, const float* scale, size_t per, void* stream) {
  if (!src || !dst || !scale) return WSDL_E_NULL;
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)WSDL_NUM_SMS * 8) blocks = (size_t)WSDL_NUM_SMS * 8;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n, scale, per);
  WSDL_LAUNCH_CHECK();
  return 0;
}
