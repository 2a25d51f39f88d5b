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
#include <stdlib.h>
#include <string.h>

#include "pairwise.cuh"

namespace wsdl {

constexpr int PW_TW = 32;       // tile width  (one lane per column)
constexpr int PW_TH = 32;       // tile height
constexpr int PW_THREADS = 256; // 8 warps, warp w owns rows w, w+8, w+16, w+24
constexpr int PW_MAXPAD = 3;    // window <= 7

// M_axis(a, e; n) = sum_{d=-pad..pad} [reflect(a+d) == a+e] * s1(d); 0 when a or a+e is outside.
__device__ float axis_multiplicity(int a, int e, int n, int pad, float inv_2ss) {
  if (a < 0 || a >= n || a + e < 0 || a + e >= n) return 0.f;
  float m = 0.f;
  for (int d = -pad; d <= pad; ++d)
    if (reflect_idx(a + d, n) == a + e) m += (inv_2ss != 0.f) ? expf(-(float)(d * d) * inv_2ss) : 1.f;
  return m;
}

// The last CTA of the launch adds the per-tile partial sums in a fixed order (deterministic), in double.
template <int BAR_ID = 0>  // 0: __syncthreads(); else a named barrier over the first PW_THREADS threads
__device__ __forceinline__ void pw_final_reduce(const PwParams& P, int tiles) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    if (BAR_ID == 0)
      __syncthreads();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(BAR_ID), "r"(PW_THREADS) : "memory");
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < PW_THREADS / 32; ++i) t += s_d[i];
      P.loss_out[0] = (float)(t * P.kappa);
    }
  }
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
  if (tid == 0) *P.ticket = 0u;  // every CTA has checked in: leave the ticket ready for the next launch
  pw_final_reduce(P, tiles);
}


// ==========================================================================================
// Fast path (window 5, C <= 4): thread = 4 horizontally adjacent pixels, 128-bit shared loads.
//
// The tile is staged WITH its reflect-padded halo, so the main loop is a plain 24-tap gather with no
// border logic at all:  S_all(z) = sum_d k(z,z+d) s(d) (p(z) - p~(z+d)),  loss += k s |p(z)-p~(z+d)|^2.
// Treating every padded position as its own variable, the true gradient is
//     g(z) = 2 kappa [ 2 S_all(z) + corr(z) ],
//     corr(z) = - sum_{d: z+d in halo} k s (p(z) - p~(z+d))          (incoming edges exist only from real pixels)
//               + sum_{u in r^-1(z), u != z} sum_{d: u+d real} k(u,u+d) s(d) (p(z) - p(u+d))   (edges into z's mirror images)
// corr is non-zero only for pixels within `pad` of the border (7% at 224x224); border CTAs compute it in a
// compacted pre-pass (lanes run along the border, not across it) and park it in shared memory.
// ==========================================================================================
constexpr int PF_PAD = 2;
constexpr int PF_SW = PW_TW + 2 * PF_PAD;  // 36 floats: 144 B rows keep 16-byte alignment of every 4-pixel run
constexpr int PF_SH = PW_TH + 2 * PF_PAD;
constexpr int PF_PLANE = PF_SH * PF_SW;

template <int C>
__device__ __forceinline__ void pf_corr_item(const PwParams& P, const float* s_img, const float* s_p, float* s_corr,
                                            int x0, int y0, int zy, int zx) {
  const int H = P.H, W = P.W;
  const int so = (zy - y0 + PF_PAD) * PF_SW + (zx - x0 + PF_PAD);
  const float i0 = s_img[so], i1 = s_img[PF_PLANE + so], i2 = s_img[2 * PF_PLANE + so];
  float pz[C], acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) pz[c] = s_p[c * PF_PLANE + so], acc[c] = 0.f;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      // a/b = 0: the pixel itself; 1: its mirror image across the low border; 2: across the high border
      const bool va = a == 0 || (a == 1 ? (zy >= 1 && zy <= PF_PAD) : (zy >= H - 1 - PF_PAD && zy <= H - 2));
      const bool vb = b == 0 || (b == 1 ? (zx >= 1 && zx <= PF_PAD) : (zx >= W - 1 - PF_PAD && zx <= W - 2));
      if (!va || !vb) continue;
      const int cy = a == 0 ? zy : (a == 1 ? -zy : 2 * (H - 1) - zy);
      const int cx = b == 0 ? zx : (b == 1 ? -zx : 2 * (W - 1) - zx);
      const bool self = (a == 0 && b == 0);
#pragma unroll 1
      for (int dy = -PF_PAD; dy <= PF_PAD; ++dy)
#pragma unroll 1
        for (int dx = -PF_PAD; dx <= PF_PAD; ++dx) {
          if (dx == 0 && dy == 0) continue;
          const int ny = cy + dy, nx = cx + dx;
          const bool real = (ny >= 0 && ny < H && nx >= 0 && nx < W);
          if (self == real) continue;  // self: halo neighbours only; mirror image: real neighbours only
          const int sn = (ny - y0 + PF_PAD) * PF_SW + (nx - x0 + PF_PAD);
          const float d0 = i0 - s_img[sn], d1 = i1 - s_img[PF_PLANE + sn], d2 = i2 - s_img[2 * PF_PLANE + sn];
          const float dist = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
          float k = ex2_approx(fmaf(dist, P.kc, (float)(dx * dx + dy * dy) * P.ks_unit));
          k = self ? -k : k;
#pragma unroll
          for (int c = 0; c < C; ++c) acc[c] = fmaf(k, pz[c] - s_p[c * PF_PLANE + sn], acc[c]);
        }
    }
  const int to = (zy - y0) * PW_TW + (zx - x0);
#pragma unroll
  for (int c = 0; c < C; ++c) s_corr[c * PW_TW * PW_TH + to] = acc[c];
}

template <int C>
__global__ void __launch_bounds__(PW_THREADS, (C <= 2 ? 3 : 2)) pairwise_fast_kernel(const __grid_constant__ PwParams P) {
  extern __shared__ __align__(16) float smem[];
  float* s_img = smem;                    // [3][PF_SH][PF_SW], reflect-padded
  float* s_p = s_img + 3 * PF_PLANE;      // [C][PF_SH][PF_SW]
  float* s_corr = s_p + C * PF_PLANE;     // [C][TH][TW], border CTAs only
  __shared__ float s_red[PW_THREADS / 32];
  __shared__ int s_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile_x = blockIdx.x, tile_y = blockIdx.y, b = blockIdx.z;
  const int x0 = tile_x * PW_TW, y0 = tile_y * PW_TH;
  const int H = P.H, W = P.W;
  const size_t plane = (size_t)H * W;
  const float* img = P.images + (size_t)b * 3 * plane;
  const float* val = P.values + (size_t)b * C * plane;

  // ---- stage tile + reflect halo; softmax on the way in.  All global loads of a thread are issued
  //      before the first use so that one memory latency covers the whole tile. ----
  {
    constexpr int NPOS = PF_SH * PF_SW;
    constexpr int NIT = (NPOS + PW_THREADS - 1) / PW_THREADS;
    float vi[NIT][3], vv[NIT][C];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = tid + it * PW_THREADS;
      const int ty = i / PF_SW, tx = i - ty * PF_SW;
      int gy = y0 - PF_PAD + ty, gx = x0 - PF_PAD + tx;
      gy = gy < 0 ? -gy : gy;
      gy = gy >= H ? 2 * (H - 1) - gy : gy;
      gx = gx < 0 ? -gx : gx;
      gx = gx >= W ? 2 * (W - 1) - gx : gx;
      gy = min(max(gy, 0), H - 1);  // positions past a ragged tile's halo: any valid address, never used
      gx = min(max(gx, 0), W - 1);
      const size_t o = (size_t)gy * W + gx;
      if (i < NPOS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) vi[it][c] = __ldg(img + c * plane + o);
#pragma unroll
        for (int c = 0; c < C; ++c) vv[it][c] = __ldg(val + c * plane + o);
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = tid + it * PW_THREADS;
      if (i < NPOS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) s_img[c * PF_PLANE + i] = vi[it][c];
        float v[C];
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = vv[it][c];
        if (P.inner_softmax) {
          float m = v[0];
#pragma unroll
          for (int c = 1; c < C; ++c) m = fmaxf(m, v[c]);
          float s = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            v[c] = ex2_approx((v[c] - m) * LOG2E);
            s += v[c];
          }
          const float inv = __fdiv_rn(1.f, s);
#pragma unroll
          for (int c = 0; c < C; ++c) v[c] *= inv;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) s_p[c * PF_PLANE + i] = v[c];
      }
    }
  }
  __syncthreads();

  const float ksx1 = P.ks_unit, ksx2 = 4.f * P.ks_unit;  // log2 of the spatial Gaussian along x

  // ---- border CTAs: correction pre-pass over the pixels within PF_PAD of an image border ----
  const int xe = min(x0 + PW_TW, W), ye = min(y0 + PW_TH, H);  // tile extent inside the image
  const bool border = (x0 <= PF_PAD) || (y0 <= PF_PAD) || (xe >= W - PF_PAD) || (ye >= H - PF_PAD);
  if (border) {
    for (int i = tid; i < C * PW_TW * PW_TH; i += PW_THREADS) s_corr[i] = 0.f;
    __syncthreads();
    const int tw = xe - x0, th = ye - y0;
    // band columns inside the tile: [x0, xl) and [xr, xe);  band rows: [y0, yl) and [yr, ye)
    const int xl = min(max(PF_PAD + 1, x0), xe), xr = max(min(W - 1 - PF_PAD, xe), xl);
    const int yl = min(max(PF_PAD + 1, y0), ye), yr = max(min(H - 1 - PF_PAD, ye), yl);
    const int ncol = (xl - x0) + (xe - xr), nrow = (yl - y0) + (ye - yr);
    for (int i = tid; i < ncol * th; i += PW_THREADS) {  // lanes run down the band columns
      const int j = i / th, ty = i - j * th;
      const int zx = j < (xl - x0) ? x0 + j : xr + (j - (xl - x0));
      pf_corr_item<C>(P, s_img, s_p, s_corr, x0, y0, y0 + ty, zx);
    }
    const int wmid = xr - xl;  // columns not already covered above
    for (int i = tid; i < nrow * wmid; i += PW_THREADS) {  // lanes run along the band rows
      const int j = i / wmid, tx = i - j * wmid;
      const int zy = j < (yl - y0) ? y0 + j : yr + (j - (yl - y0));
      pf_corr_item<C>(P, s_img, s_p, s_corr, x0, y0, zy, xl + tx);
    }
    (void)tw;
    __syncthreads();
  }

  // ---- main pass: lane = (run of 4 pixels, row): 8 runs x 4 rows per warp, 8 warps = 32 rows ----
  const int run = lane & 7, row = warp * 4 + (lane >> 3);
  const int gy = y0 + row, gx = x0 + run * 4;
  float g[4][C];
  float loss_acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < C; ++c) g[j][c] = 0.f;
  float ci[3][4], cp[C][4];  // centre pixels
  {
    const int so = (row + PF_PAD) * PF_SW + run * 4 + PF_PAD;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ci[c][j] = s_img[c * PF_PLANE + so + j];
#pragma unroll
      for (int c = 0; c < C; ++c) cp[c][j] = s_p[c * PF_PLANE + so + j];
    }
  }
#pragma unroll 1
  for (int ey = -PF_PAD; ey <= PF_PAD; ++ey) {
    const int so = (row + PF_PAD + ey) * PF_SW + run * 4;  // window columns gx-2 .. gx+5
    const float ksy = (float)(ey * ey) * P.ks_unit;
    const float kse[5] = {ksy + ksx2, ksy + ksx1, ksy, ksy + ksx1, ksy + ksx2};
    float wi[3][8], wp[C][8];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(s_img + c * PF_PLANE + so);
      const float4 d = *reinterpret_cast<const float4*>(s_img + c * PF_PLANE + so + 4);
      wi[c][0] = a.x, wi[c][1] = a.y, wi[c][2] = a.z, wi[c][3] = a.w;
      wi[c][4] = d.x, wi[c][5] = d.y, wi[c][6] = d.z, wi[c][7] = d.w;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(s_p + c * PF_PLANE + so);
      const float4 d = *reinterpret_cast<const float4*>(s_p + c * PF_PLANE + so + 4);
      wp[c][0] = a.x, wp[c][1] = a.y, wp[c][2] = a.z, wp[c][3] = a.w;
      wp[c][4] = d.x, wp[c][5] = d.y, wp[c][6] = d.z, wp[c][7] = d.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int ex = -PF_PAD; ex <= PF_PAD; ++ex) {
        const int n = j + PF_PAD + ex;  // window index of the neighbour (the centre tap adds exactly 0)
        const float d0 = ci[0][j] - wi[0][n], d1 = ci[1][j] - wi[1][n], d2 = ci[2][j] - wi[2][n];
        const float dist = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
        const float k = ex2_approx(fmaf(dist, P.kc, kse[ex + PF_PAD]));
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float dp = cp[c][j] - wp[c][n];
          sq = fmaf(dp, dp, sq);
          g[j][c] = fmaf(k, dp, g[j][c]);
        }
        loss_acc = fmaf(k, sq, loss_acc);
      }
    }
  }

  // ---- epilogue: g = 2 kappa (2 S_all + corr), softmax backward, store; loss partial ----
  const float scale_g = (float)(2.0 * P.kappa) * (P.grad_out ? __ldg(P.grad_out + (P.per_image ? b : 0)) : 1.f);
  float lsum = 0.f;
  if (gy < H) {
    float out[C][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        gg[c] = 2.f * g[j][c];
        if (border) gg[c] += s_corr[c * PW_TW * PW_TH + row * PW_TW + run * 4 + j];
      }
      if (P.inner_softmax) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dot = fmaf(cp[c][j], gg[c], dot);
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][j] = scale_g * cp[c][j] * (gg[c] - dot);
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) out[c][j] = scale_g * gg[c];
      }
    }
    if (gx + 3 < W) lsum = loss_acc;  // all four pixels real (the common case); ragged runs are redone below
    if (P.grad_values) {
      float* go = P.grad_values + (size_t)b * C * plane + (size_t)gy * W + gx;
      const bool vec = (gx + 3 < W) && ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.grad_values) & 15) == 0);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (vec) {
          *reinterpret_cast<float4*>(go + c * plane) = make_float4(out[c][0], out[c][1], out[c][2], out[c][3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gx + j < W) go[c * plane + j] = out[c][j];
        }
      }
    }
  }
  if (gy < H && gx < W && gx + 3 >= W) {
    // ragged run at the right edge: loss of the real pixels only (the tile's extra columns hold clamped data)
    const int so = (row + PF_PAD) * PF_SW + run * 4 + PF_PAD;
    for (int j = 0; gx + j < W; ++j)
      for (int ey = -PF_PAD; ey <= PF_PAD; ++ey)
        for (int ex = -PF_PAD; ex <= PF_PAD; ++ex) {
          if (ex == 0 && ey == 0) continue;
          const int sn = so + j + ey * PF_SW + ex;
          const float d0 = s_img[so + j] - s_img[sn], d1 = s_img[PF_PLANE + so + j] - s_img[PF_PLANE + sn],
                      d2 = s_img[2 * PF_PLANE + so + j] - s_img[2 * PF_PLANE + sn];
          const float k = ex2_approx(fmaf(fmaf(d2, d2, fmaf(d1, d1, d0 * d0)), P.kc, (float)(ex * ex + ey * ey) * P.ks_unit));
          float sq = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float dp = s_p[c * PF_PLANE + so + j] - s_p[c * PF_PLANE + sn];
            sq = fmaf(dp, dp, sq);
          }
          lsum = fmaf(k, sq, lsum);
        }
  }

  lsum = warp_sum(lsum);
  if (lane == 0) s_red[warp] = lsum;
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
  if (tid == 0) *P.ticket = 0u;  // every CTA has checked in: leave the ticket ready for the next launch
  pw_final_reduce(P, tiles);
}

template <int C>
static int pf_launch(const PwParams& P, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)(3 + C) * PF_PLANE + (size_t)C * PW_TW * PW_TH);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_fast_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(P.tiles_x, P.tiles_y, P.B);
  pairwise_fast_kernel<C><<<grid, PW_THREADS, smem, s>>>(P);
  WSDL_LAUNCH_CHECK();
  return 0;
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
  size_t floats = (size_t)B * tiles;  // one partial per 32x32 tile (generic / fast kernels)
  const size_t sym = ps_workspace_floats(B, H, W);
  if (sym > floats) floats = sym;
  // ... and, behind them, the result slots of the pair-symmetric kernels (one word per CTA): a region no other kernel writes
  return 512 + pw_align(floats * sizeof(float)) + pw_align(sym * sizeof(unsigned));
}

static unsigned* pw_sym_slots(uintptr_t ws, int B, int H, int W) {
  const size_t sym_bytes = pw_align(ps_workspace_floats(B, H, W) * sizeof(unsigned));
  return reinterpret_cast<unsigned*>(ws + (wsdl_pairwise_workspace_bytes(B, H, W) - sym_bytes));
}

static int pairwise_fwd_bwd_impl(const float* values, const float* images, int B, int C, int H, int W, int window,
                                 float sigma_color, float sigma_space, int inner_softmax, int divide_by_c,
                                 int per_image_loss, const float* grad_out, float* loss_out, float* grad_values,
                                 void* workspace, size_t workspace_bytes, void* stream, bool prepared) {
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
  P.sym_slots = pw_sym_slots(ws, B, H, W);
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
  if (!prepared) {  // every kernel leaves its ticket at 0 / its slots empty again: a prepared workspace needs no per-call memset
    cudaError_t e = cudaMemsetAsync(P.ticket, 0, 8, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(P.sym_slots, 0xff, ps_workspace_floats(B, H, W) * sizeof(unsigned), s);
    if (e != cudaSuccess) return (int)e;
  }
  if (pad == 2 && H > 2 * PF_PAD && W > 2 * PF_PAD) {
    static const int no_sym = WSDL_TUNE_INT("WSDL_PAIRWISE_NO_SYM", 0);
    if (!no_sym) {  // pair-symmetric kernels: the hot configurations (window 5, C <= 2)
      const int rc = ps_launch(P, s);  // one tile per CTA (pairwise_sym.cu)
      if (rc != 1) return rc;
    }
    switch (C) {
      case 1: return pf_launch<1>(P, s);
      case 2: return pf_launch<2>(P, s);
      case 3: return pf_launch<3>(P, s);
      case 4: return pf_launch<4>(P, s);
      default: break;
    }
  }
  if (pad == 2 && C == 2) return pw_launch<2, 2>(P, s);
  return pw_launch<0, 0>(P, s);
}

extern "C" int wsdl_pairwise_fwd_bwd(const float* values, const float* images, int B, int C, int H, int W,
                                     int window, float sigma_color, float sigma_space, int inner_softmax,
                                     int divide_by_c, int per_image_loss, const float* grad_out, float* loss_out,
                                     float* grad_values, void* workspace, size_t workspace_bytes, void* stream) {
  return pairwise_fwd_bwd_impl(values, images, B, C, H, W, window, sigma_color, sigma_space, inner_softmax, divide_by_c,
                               per_image_loss, grad_out, loss_out, grad_values, workspace, workspace_bytes, stream, false);
}

// A prepared workspace: tickets / control words zero, everything behind them "empty" (all ones -- the state the fused
// launch's per-CTA result slots are polled for and left in; the other kernels write their partials before they read them).
extern "C" int wsdl_pairwise_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace) return WSDL_E_NULL;
  if (workspace_bytes < 512) return WSDL_E_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(workspace, 0, 512, (cudaStream_t)stream);
  if (e == cudaSuccess && workspace_bytes > 512)
    e = cudaMemsetAsync(reinterpret_cast<unsigned char*>(workspace) + 512, 0xff, workspace_bytes - 512, (cudaStream_t)stream);
  return (int)e;
}

extern "C" int wsdl_pairwise_fwd_bwd_prepared(const float* values, const float* images, int B, int C, int H, int W,
                                              int window, float sigma_color, float sigma_space, int inner_softmax,
                                              int divide_by_c, int per_image_loss, const float* grad_out,
                                              float* loss_out, float* grad_values, void* workspace,
                                              size_t workspace_bytes, void* stream) {
  return pairwise_fwd_bwd_impl(values, images, B, C, H, W, window, sigma_color, sigma_space, inner_softmax, divide_by_c,
                               per_image_loss, grad_out, loss_out, grad_values, workspace, workspace_bytes, stream, true);
}

static size_t sg_bytes(int B, int H, int W) { return 512 + pw_align(sg_workspace_floats(B, H, W) * sizeof(float)); }

// result slots of the fused launch (one 64-bit word per CTA), a region of their own at the end of the workspace: no other
// kernel that may share a prepared workspace ever writes there
static size_t dual_slot_bytes(int B, int H, int W) { return pw_align(ps_workspace_floats(B, H, W) * sizeof(unsigned long long)); }

extern "C" size_t wsdl_pairwise_dual_workspace_bytes(int B, int H, int W) {
  const size_t one = wsdl_pairwise_workspace_bytes(B, H, W);
  if (!one) return 0;
  const size_t sg = sg_bytes(B, H, W);
  return (2 * one > sg ? 2 * one : sg) + 256 + dual_slot_bytes(B, H, W);
}

extern "C" size_t wsdl_weak_loss_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  return sg_bytes(B, H, W) + 256;
}

extern "C" int wsdl_pairwise_dual_fwd_bwd(const float* logits, const float* images, int B, int H, int W, int window,
                                          float sigma_cut, float sigma_bnd, float sigma_space,
                                          const float* grad_out_cut, const float* grad_out_bnd, float* loss_cut,
                                          float* loss_bnd, float* grad_logits, void* workspace, size_t workspace_bytes,
                                          int prepared, void* stream) {
  if (!logits || !images || !loss_cut || !loss_bnd || !workspace) return WSDL_E_NULL;
  if (B < 1 || B > 65535 || H < 6 || W < 6 || window != 5) return WSDL_E_SHAPE;  // the pair-symmetric kernel's shapes
  if (!(sigma_cut > 0.f) || !(sigma_bnd > 0.f)) return WSDL_E_ARG;
  if (((uintptr_t)logits % 4) || ((uintptr_t)images % 4) || ((uintptr_t)loss_cut % 4) || ((uintptr_t)loss_bnd % 4) ||
      (grad_logits && ((uintptr_t)grad_logits % 4)))
    return WSDL_E_ALIGN;
  const size_t one = wsdl_pairwise_workspace_bytes(B, H, W);
  if (workspace_bytes < wsdl_pairwise_dual_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  // Two regularisers only, f32: the one-tile-per-CTA kernel (pairwise_sym.cu) is the faster one when launches overlap
  // (22 vs 30 us per launch on configs[1]; 37 vs 41 us for a lone launch).  The tile-streaming kernel
  // (pairwise_stream.cu) is the one that also takes labels, bf16 logits and u8 images: wsdl_weak_loss_fwd_bwd.
  static const int no_stream = !WSDL_TUNE_INT("WSDL_PAIRWISE_DUAL_STREAM", 0);
  if (!no_stream) {
    SgLaunch L;
    memset(&L, 0, sizeof(L));
    L.logits = logits, L.images = images, L.grad = grad_logits, L.go_cut = grad_out_cut, L.go_bnd = grad_out_bnd;
    L.loss_cut = loss_cut, L.loss_bnd = loss_bnd;
    const uintptr_t w0 = ((uintptr_t)workspace + 255) / 256 * 256;
    L.ctrl = reinterpret_cast<unsigned*>(w0);
    L.partial = reinterpret_cast<float*>(w0 + 256);
    L.B = B, L.H = H, L.W = W;
    L.logit_dtype = WSDL_F32, L.image_dtype = WSDL_F32, L.grad_dtype = WSDL_F32;
    L.sigma_cut = sigma_cut, L.sigma_bnd = sigma_bnd, L.sigma_space = sigma_space;
    if (!prepared) {
      cudaError_t e = cudaMemsetAsync(L.ctrl, 0, 8, s);
      if (e != cudaSuccess) return (int)e;
    }
    const int rc = sg_launch(L, s);
    if (rc != 1) return rc;
  }
  PwParams P;
  P.values = logits;
  P.images = images;
  P.grad_out = grad_out_cut;
  P.loss_out = loss_cut;
  P.grad_values = grad_logits;
  uintptr_t ws = ((uintptr_t)workspace + 255) / 256 * 256;
  P.ticket = reinterpret_cast<unsigned*>(ws);
  P.partial = reinterpret_cast<float*>(ws + 256);
  P.sym_slots = nullptr;
  const size_t slot_bytes = dual_slot_bytes(B, H, W);
  unsigned long long* slots = reinterpret_cast<unsigned long long*>(ws + (wsdl_pairwise_dual_workspace_bytes(B, H, W) - 256 - slot_bytes));
  P.B = B, P.C = 2, P.H = H, P.W = W, P.pad = 2;
  P.tiles_x = (W + PW_TW - 1) / PW_TW;
  P.tiles_y = (H + PW_TH - 1) / PW_TH;
  P.inner_softmax = 1;
  P.per_image = 0;
  P.kc = -LOG2E / (2.f * sigma_cut * sigma_cut);
  P.inv_2ss = 0.f;
  P.ks_unit = 0.f;
  P.kappa = 1.0 / (24.0 * (double)B * (double)H * (double)W * 2.0);
  if (!prepared && no_stream) {
    cudaError_t e = cudaMemsetAsync(slots, 0xff, slot_bytes, s);
    if (e != cudaSuccess) return (int)e;
  }
  const int rc = ps_launch_dual(P, sigma_cut, sigma_bnd, sigma_space, grad_out_bnd, loss_bnd, slots, s);
  return rc == 1 ? WSDL_E_SHAPE : rc;
}

__global__ void count_valid_labels_kernel(const void* labels, int dtype, size_t n, long long ignore_index,
                                          unsigned long long* scratch, unsigned* done, float* inv_count) {
  unsigned long long c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const long long v = dtype == WSDL_U8 ? (long long)__ldg(reinterpret_cast<const unsigned char*>(labels) + i)
                                         : __ldg(reinterpret_cast<const long long*>(labels) + i);
    c += v != ignore_index;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  __shared__ unsigned long long s_c[8];
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_c[i];
    atomicAdd(scratch, t);
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {  // last block: exact integer count -> reciprocal, as torch divides
      __threadfence();
      const unsigned long long total = *reinterpret_cast<volatile unsigned long long*>(scratch);
      inv_count[0] = __fdiv_rn(1.f, (float)total);
    }
  }
}

extern "C" int wsdl_count_valid_labels(const void* labels, int labels_dtype, size_t n, long long ignore_index,
                                       unsigned long long* scratch, float* inv_count, void* stream) {
  if (!labels || !scratch || !inv_count) return WSDL_E_NULL;
  if (labels_dtype != WSDL_U8 && labels_dtype != WSDL_I64) return WSDL_E_DTYPE;
  if (((uintptr_t)scratch % 8) || (labels_dtype == WSDL_I64 && ((uintptr_t)labels % 8))) return WSDL_E_ALIGN;
  if (n == 0) return WSDL_E_SHAPE;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, 16, s);  // the count and the block check-in word behind it
  if (e != cudaSuccess) return (int)e;
  size_t blocks = (n + 256 * 8 - 1) / (256 * 8);
  if (blocks > (size_t)WSDL_NUM_SMS * 4) blocks = (size_t)WSDL_NUM_SMS * 4;
  count_valid_labels_kernel<<<(unsigned)blocks, 256, 0, s>>>(labels, labels_dtype, n, ignore_index, scratch,
                                                             reinterpret_cast<unsigned*>(scratch + 1), inv_count);
  WSDL_LAUNCH_CHECK();
  return 0;
}

extern "C" int wsdl_weak_loss_fwd_bwd(const void* logits, int logits_dtype, const void* images, int images_dtype,
                                      const void* labels, int labels_dtype, long long ignore_index, int B, int H, int W,
                                      int window, float sigma_cut, float sigma_bnd, float sigma_space, float lam_ce,
                                      const float* ce_inv_count, const float* grad_out_cut, const float* grad_out_bnd,
                                      float* loss_ce, float* loss_cut, float* loss_bnd, float* loss_total,
                                      void* grad_logits, int grad_dtype, void* workspace, size_t workspace_bytes, int prepared, void* stream) {
  if (!logits || !images || !loss_cut || !loss_bnd || !workspace || (labels && !loss_ce)) return WSDL_E_NULL;
  if (B < 1 || B > 65535 || H < 6 || W < 6 || window != 5) return WSDL_E_SHAPE;
  if (!(sigma_cut > 0.f) || !(sigma_bnd > 0.f)) return WSDL_E_ARG;
  if (logits_dtype != WSDL_F32 && logits_dtype != WSDL_BF16) return WSDL_E_DTYPE;
  if (images_dtype != WSDL_F32 && images_dtype != WSDL_U8) return WSDL_E_DTYPE;
  if (labels && labels_dtype != WSDL_U8 && labels_dtype != WSDL_I64) return WSDL_E_DTYPE;
  if (grad_logits && grad_dtype != WSDL_F32 && grad_dtype != WSDL_BF16) return WSDL_E_DTYPE;
  if (((uintptr_t)loss_cut % 4) || ((uintptr_t)loss_bnd % 4) || (loss_ce && ((uintptr_t)loss_ce % 4)) ||
      (labels && ((uintptr_t)labels % (labels_dtype == WSDL_I64 ? 16 : 2))))  // labels are read two at a time
    return WSDL_E_ALIGN;
  if (workspace_bytes < wsdl_weak_loss_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  SgLaunch L;
  memset(&L, 0, sizeof(L));
  L.logits = logits, L.images = images, L.labels = labels, L.grad = grad_logits;
  L.go_cut = grad_out_cut, L.go_bnd = grad_out_bnd, L.ce_inv_n_dev = ce_inv_count;
  L.loss_cut = loss_cut, L.loss_bnd = loss_bnd, L.loss_ce = loss_ce, L.loss_total = loss_total;
  const uintptr_t w0 = ((uintptr_t)workspace + 255) / 256 * 256;
  L.ctrl = reinterpret_cast<unsigned*>(w0);
  L.partial = reinterpret_cast<float*>(w0 + 256);
  L.ignore_index = ignore_index;
  L.B = B, L.H = H, L.W = W;
  L.logit_dtype = logits_dtype, L.image_dtype = images_dtype, L.label_dtype = labels_dtype, L.grad_dtype = grad_dtype;
  L.sigma_cut = sigma_cut, L.sigma_bnd = sigma_bnd, L.sigma_space = sigma_space, L.lam_ce = lam_ce;
  if (!prepared) {
    cudaError_t e = cudaMemsetAsync(L.ctrl, 0, 8, s);
    if (e != cudaSuccess) return (int)e;
  }
  const int rc = sg_launch(L, s);
  return rc == 1 ? WSDL_E_SHAPE : rc;
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
