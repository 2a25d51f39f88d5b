// Pair-symmetric cut / boundary loss, persistent warp-specialised pipeline (window 5, C <= 2), sm_100a.
//
// Same arithmetic as pairwise_sym.cu (see its header for the pair symmetry, the Euler loss, the border
// multiplicities and the sentinel colour; reference: LocalNormalizedCutLoss.forward, TraditionalModel/
// AlternatingDirectionCutLoss.py:71-105; ConstrainToBoundaryLossSingle.forward,
// AlternatingDirectionBoundaryLoss.py:20-70).  What changes is WHEN things happen.  Per-warp traces of the one-tile-
// per-CTA kernel showed a CTA marching for only ~40 % of its residency: the rest is waiting for its tile, converting
// it and finishing segment heads, and the four CTAs of an SM do these in lock step.  Here a CTA is persistent
// (2 per SM) and split into two roles that overlap on different tiles:
//
//   M  warps 0-3  march tile j (the 12 forward pairs per pixel), add the carries into the segment heads, finish
//                 the band-column pixels, then hand the tile over;
//   H  warps 4-7  keep two tile buffers fed: TMA load of tile j+1 (issued as soon as M releases that buffer's
//                 planes), in-place conversion (image * sqrt(-kc), logits -> probabilities, sentinel outside the
//                 image), and the tail of tile j-1 (emit the first two rows of segments 1..7 from the head arrays).
//
// Hand-offs are mbarriers: full[b] (TMA landed) -> ready[b] (converted) -> marched[b] (planes free, heads
// complete) -> taildone[b] (head arrays free).  Loss: two partials per tile (M and H), fire-and-forget ticket, the
// H warps of the last CTA add them per image in a fixed order in double.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "pairwise.cuh"

namespace wsdl {
namespace pp {

constexpr int PP_TW = 60;     // owned columns per tile
constexpr int PP_PITCH = 68;  // smem row: image x0-4 .. x0+63 (2 pad + 2 halo | 60 | 2 halo + 2 pad)
constexpr int PP_Q = PP_PITCH / 4;
constexpr int PP_SEGS = 16;   // row segments per tile (two rows each), one per half warp of the M group
constexpr int PP_M_THREADS = 256;  // marching warps 0-7
constexpr int PP_H_THREADS = 128;  // loading / converting warps 8-11
constexpr int PP_THREADS = PP_M_THREADS + PP_H_THREADS;
constexpr int PP_CTAS_PER_SM = 2;
constexpr int PP_M_REGS = 96, PP_H_REGS = 48;  // setmaxnreg: 256 * 96 + 128 * 48 = 384 * 80 (the launch allocation)
constexpr float PP_SENTINEL = 1e15f;  // staged colour (channel 0) of positions outside the image: 2^-(1e30) = 0

struct PpParams {
  PwParams p;
  int n_x;          // column tiles per image
  int nb;           // row blocks per column tile
  int n_tiles;      // nb * n_x * B, tile id = (image * n_x + column tile) * nb + row block
  int S;            // rows per segment of this launch
  int use_tma;
  int vec2_ok;      // W % 2 == 0 and 8-byte aligned gradient: float2 stores
  float img_scale;  // sqrt(-kc)
  float g1, g4;     // gamma, gamma^4
  float l32, l1g;   // log2(3/2), log2(1 + gamma^4)
};

struct PpKs {  // exponent offsets of one row step: spatial term + row-border multiplicity
  float a1, a4;      // partner in the same row, dx^2 = 1, 4
  float b0, b1, b4;  // one row down,  dx^2 = 0, 1, 4
  float c0, c1, c4;  // two rows down
};

template <int C>
struct PpWin {  // an 8-column window of one staged row: columns 4*strip .. 4*strip+7 of the smem row
  float i[3][8];
  float p[C][8];
};

template <int C, int PL>
__device__ __forceinline__ void pp_load(PpWin<C>& w, const float* s_img, const float* s_p, int off) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(s_img + c * PL + off);
    const float4 b = *reinterpret_cast<const float4*>(s_img + c * PL + off + 4);
    w.i[c][0] = a.x, w.i[c][1] = a.y, w.i[c][2] = a.z, w.i[c][3] = a.w;
    w.i[c][4] = b.x, w.i[c][5] = b.y, w.i[c][6] = b.z, w.i[c][7] = b.w;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(s_p + c * PL + off);
    const float4 b = *reinterpret_cast<const float4*>(s_p + c * PL + off + 4);
    w.p[c][0] = a.x, w.p[c][1] = a.y, w.p[c][2] = a.z, w.p[c][3] = a.w;
    w.p[c][4] = b.x, w.p[c][5] = b.y, w.p[c][6] = b.z, w.p[c][7] = b.w;
  }
}

// one unordered pair: k = 2^(ks - |I'(a) - I'(b)|^2);  G(a) += k (p(a) - p(b));  G(b) -= k (p(a) - p(b))
template <int C>
__device__ __forceinline__ void pp_pair(float (&ga)[C], float (&gb)[C], const PpWin<C>& a, int ia, const PpWin<C>& b,
                                        int ib, float ks) {
  const float d0 = a.i[0][ia] - b.i[0][ib], d1 = a.i[1][ia] - b.i[1][ib], d2 = a.i[2][ia] - b.i[2][ib];
  const float k = ex2_approx(fmaf(-d2, d2, fmaf(-d1, d1, fmaf(-d0, d0, ks))));
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float dp = a.p[c][ia] - b.p[c][ib];
    ga[c] = fmaf(k, dp, ga[c]);
    gb[c] = fmaf(-k, dp, gb[c]);
  }
}

// All 12 forward pairs of the 4 centres of row t (accumulator X), partners in rows t (X), t+1 (Y), t+2 (Z).
template <int C, int PL>
__device__ __forceinline__ void pp_step(float (&X)[8][C], float (&Y)[8][C], float (&Z)[8][C], float (&pc)[4][C],
                                        const float* s_img, const float* s_p, int off, const PpKs& ks) {
  PpWin<C> c;
  pp_load<C, PL>(c, s_img, s_p, off);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int cc = 0; cc < C; ++cc) pc[j][cc] = c.p[cc][2 + j];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pp_pair<C>(X[2 + j], X[3 + j], c, 2 + j, c, 3 + j, ks.a1);
    pp_pair<C>(X[2 + j], X[4 + j], c, 2 + j, c, 4 + j, ks.a4);
  }
  PpWin<C> n;
  pp_load<C, PL>(n, s_img, s_p, off + PP_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx)
      pp_pair<C>(X[2 + j], Y[2 + j + dx], c, 2 + j, n, 2 + j + dx, dx == 0 ? ks.b0 : (dx * dx == 1 ? ks.b1 : ks.b4));
  pp_load<C, PL>(n, s_img, s_p, off + 2 * PP_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx)
      pp_pair<C>(X[2 + j], Z[2 + j + dx], c, 2 + j, n, 2 + j + dx, dx == 0 ? ks.c0 : (dx * dx == 1 ? ks.c1 : ks.c4));
}

// A finished accumulator row: the two columns either side of the strip belong to the neighbouring lanes.
template <int C>
__device__ __forceinline__ void pp_exchange(const float (&X)[8][C], float (&own)[4][C], int strip) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float r0 = __shfl_down_sync(0xffffffffu, X[0][c], 1), r1 = __shfl_down_sync(0xffffffffu, X[1][c], 1);
    float l0 = __shfl_up_sync(0xffffffffu, X[6][c], 1), l1 = __shfl_up_sync(0xffffffffu, X[7][c], 1);
    if (strip == 15) r0 = 0.f, r1 = 0.f;
    if (strip == 0) l0 = 0.f, l1 = 0.f;
    own[0][c] = X[2][c] + l0;
    own[1][c] = X[3][c] + l1;
    own[2][c] = X[4][c] + r0;
    own[3][c] = X[5][c] + r1;
  }
}

struct PpBlk {
  int b, x0, ys, n, nc;
  bool xband;    // the tile owns pixels within 3 columns of the left / right image border
  float scale2;  // 4 kappa * upstream gradient: dL/dp = scale2 * G
};

// band slot of a coordinate: 0..2 for the low band, 3..5 for the high band, -1 outside (needs n >= 6)
__device__ __forceinline__ int pp_band_slot(int v, int n) { return v <= 2 ? v : (v >= n - 3 ? v - (n - 6) : -1); }

// One pixel: dL/dvalue from G (linear in G: softmax backward included), and its term of 2 kappa sum (p - 1/2) G.
template <int C, int CS, bool SOFTMAX>
__device__ __forceinline__ float pp_pixel_grad(float scale2, const float (&p)[CS], const float (&g)[CS], float (&out)[C]) {
  float l;
  if (CS != C) {  // two classes behind a softmax: p1 = 1 - p0, G1 = -G0
    const float p0 = p[0], p1 = 1.f - p0;
    l = (p0 - p1) * g[0];
    out[0] = scale2 * 2.f * p0 * p1 * g[0];
    out[C - 1] = -out[0];
  } else {
    l = 0.f;
#pragma unroll
    for (int c = 0; c < CS; ++c) l = fmaf(p[c] - 0.5f, g[c], l);
    if (SOFTMAX) {
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) dot = fmaf(p[c], g[c], dot);
#pragma unroll
      for (int c = 0; c < CS; ++c) out[c] = scale2 * p[c] * (g[c] - dot);
    } else {
#pragma unroll
      for (int c = 0; c < CS; ++c) out[c] = scale2 * g[c];
    }
  }
  return l;
}

// Gradient of 4 finished pixels of centre row t: store, and accumulate the loss term.  Pixels of the band columns
// (bandmask) are neither stored nor counted here: the band pass finishes them from their G.
template <int C, int CS, bool SOFTMAX>
__device__ __forceinline__ void pp_emit(const PpParams& Q, const PpBlk& K, int t, int strip, int okmask, int bandmask,
                                        const float (&G)[4][CS], const float (&pc)[4][CS], float& lsum) {
  const int H = Q.p.H, W = Q.p.W;
  const int y = K.ys - 2 + t;
  const int xs = K.x0 - 2 + 4 * strip;  // image column of j = 0
  const int m = okmask & ~bandmask;
  float out[C][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float o[C];
    const float l = pp_pixel_grad<C, CS, SOFTMAX>(K.scale2, pc[j], G[j], o);
#pragma unroll
    for (int c = 0; c < C; ++c) out[c][j] = o[c];
    if ((m >> j) & 1) lsum += l;
  }
  if (Q.p.grad_values) {
    const size_t plane = (size_t)H * W;
    float* go = Q.p.grad_values + (size_t)K.b * C * plane + (size_t)y * W + xs;
    if (Q.vec2_ok && bandmask == 0) {  // xs is even, so (xs, xs+1) and (xs+2, xs+3) are inside or outside W together
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (okmask & 1) *reinterpret_cast<float2*>(go + c * plane) = make_float2(out[c][0], out[c][1]);
        if (okmask & 4) *reinterpret_cast<float2*>(go + c * plane + 2) = make_float2(out[c][2], out[c][3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if ((m >> j) & 1) go[c * plane + j] = out[c][j];
    }
  }
}

// gamma^(k^2) for k = 0, 1, 2
__device__ __forceinline__ float pp_gpow(int k, float g1, float g4) { return k == 0 ? 1.f : (k == 1 ? g1 : g4); }

// W(u -> v) = sum_{d=-2..2} [reflect(u + d) == v] gamma^(d^2) for u inside [0, n), n >= 6; 0 for v outside (closed
// form of pairwise.cu's axis_multiplicity for pad 2).
__device__ __forceinline__ float pp_w1d(int u, int v, int n, float g1, float g4) {
  if (v < 0 || v >= n) return 0.f;
  const int d = abs(v - u), s = u + v, e = 2 * (n - 1) - s;
  float w = d <= 2 ? pp_gpow(d, g1, g4) : 0.f;
  if (v >= 1 && s <= 2) w += pp_gpow(s, g1, g4);      // offset -s lands on -v, which reflects to v
  if (v <= n - 2 && e <= 2) w += pp_gpow(e, g1, g4);  // offset +e lands on 2(n-1) - v
  return w;
}

// multiplicity the march applies to a pair of rows (relative to the interior 2 gamma^|d|^2): see the row step
__device__ __forceinline__ float pp_row_mult(int ya, int yb, int H, float g4) {
  const int lo = min(ya, yb), d = abs(ya - yb);
  if (d == 0) return (lo == 1 || lo == H - 2) ? 1.f + g4 : 1.f;
  return (lo == 0 || lo + d == H - 1) ? 1.5f : 1.f;
}

// xfix(a) for one band pixel a = (zy, zx): sum over its in-image window partners b of
// (true pair weight - weight applied by the march) / 2 * kc(a,b) (p(a) - p(b)).  Reads the staged tile; s_wx holds
// Wx(zx -> zx + j - 2) and Wx(zx + j - 2 -> zx), j = 0..4, for the band slot of zx.  The difference vanishes unless
// the COLUMN weights of the pair differ from the interior gamma^dx^2 (the march's row multiplicity is exactly
// (Wy + Wy') / (2 gamma^dy^2)), which happens for at most two partner columns of a band pixel; those two columns are
// evaluated for all five rows without branches (weight 0 where there is nothing to add), so the loads and the
// exponentials of the ten candidates overlap.
template <int CS, int PL>
__device__ __forceinline__ void pp_xfix_item(const PpParams& Q, const float* s_img, const float* s_p, const float* s_wx,
                                            int ys, int x0, int zy, int zx, float (&acc)[CS], float (&pz)[CS]) {
  const int H = Q.p.H;
  const float g1 = Q.g1, g4 = Q.g4;
  const int so = (zy - (ys - 2)) * PP_PITCH + (zx - (x0 - 4));
  const float i0 = s_img[so], i1 = s_img[PL + so], i2 = s_img[2 * PL + so];
#pragma unroll
  for (int c = 0; c < CS; ++c) pz[c] = s_p[c * PL + so], acc[c] = 0.f;
  // the (at most two) partner columns whose weights are not the interior ones
  int jsp[2] = {-1, -1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float f = s_wx[j], b = s_wx[5 + j], g = pp_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
    if (f != 0.f && (f != g || b != g)) {
      if (jsp[0] < 0) jsp[0] = j;
      else jsp[1] = j;
    }
  }
  // row weights: forward, backward, and what the march applies (2 gamma^dy^2 times its row multiplicity)
  float wyf[5], wyb[5], wym[5];
  if (zy >= 5 && zy <= H - 6) {  // every partner row is clear of the row bands
#pragma unroll
    for (int i = 0; i < 5; ++i) wyf[i] = wyb[i] = pp_gpow(i < 2 ? 2 - i : i - 2, g1, g4), wym[i] = 2.f * wyf[i];
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yb = zy + i - 2;
      const bool in = yb >= 0 && yb < H;
      wyf[i] = in ? pp_w1d(zy, yb, H, g1, g4) : 0.f;
      wyb[i] = in ? pp_w1d(yb, zy, H, g1, g4) : 0.f;
      wym[i] = in ? 2.f * pp_gpow(i < 2 ? 2 - i : i - 2, g1, g4) * pp_row_mult(zy, yb, H, g4) : 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = jsp[u] < 0 ? 2 : jsp[u];  // no such column: the pixel's own (weight 0 below)
    const float live = jsp[u] < 0 ? 0.f : 0.5f;
    const float wxf = s_wx[j], wxb = s_wx[5 + j], gx = pp_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float diff = live * (fmaf(wyf[i], wxf, wyb[i] * wxb) - wym[i] * gx);  // 0 for rows outside the image
      const int sn = so + (i - 2) * PP_PITCH + (j - 2);
      const float d0 = i0 - s_img[sn], d1 = i1 - s_img[PL + sn], d2 = i2 - s_img[2 * PL + sn];
      const float k = diff * ex2_approx(fmaf(-d2, d2, fmaf(-d1, d1, -d0 * d0)));
#pragma unroll
      for (int c = 0; c < CS; ++c) acc[c] = fmaf(k, pz[c] - s_p[c * PL + sn], acc[c]);
    }
  }
}

// ---- staging ----
__device__ __forceinline__ unsigned pp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Rows [0, rows) of the tile, element by element, raw values (0 outside the image): the layout a TMA load leaves.
// Work is spread over a thread group: thread `gt` of `gn`.
template <int C, int PL>
__device__ __noinline__ void pp_tile_load_slow(const PpParams& Q, const PpBlk& K, float* s_img, float* s_val, int rows,
                                               int gt, int gn) {
  const int H = Q.p.H, W = Q.p.W;
  const size_t plane = (size_t)H * W;
  const float* img = Q.p.images + (size_t)K.b * 3 * plane;
  const float* val = Q.p.values + (size_t)K.b * C * plane;
  for (int i = gt; i < rows * PP_PITCH; i += gn) {
    const int t = i / PP_PITCH, cs = i - t * PP_PITCH;
    const int y = K.ys - 2 + t, x = K.x0 - 4 + cs;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    const size_t o = in ? (size_t)y * W + x : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) s_img[c * PL + i] = in ? __ldg(img + c * plane + o) : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s_val[c * PL + i] = in ? __ldg(val + c * plane + o) : 0.f;
  }
}

// Rows [0, rows) of the tile, in place: image * sqrt(-kc) (sentinel outside the image), values -> probabilities.
template <int C, int CS, bool SOFTMAX, int PL>
__device__ __forceinline__ void pp_tile_transform(const PpParams& Q, const PpBlk& K, float* s_img, float* s_val, int rows,
                                                  int gt, int gn) {
  const int H = Q.p.H, W = Q.p.W;
  const float sc = Q.img_scale;
#pragma unroll 1
  for (int it = gt; it < rows * PP_Q; it += gn) {
    const int t = it / PP_Q, q = it - t * PP_Q;
    const int y = K.ys - 2 + t, xb = K.x0 - 4 + 4 * q;
    const bool row_in = y >= 0 && y < H;
    const int so = it * 4;  // PP_PITCH == 4 * PP_Q
    float4 v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = *reinterpret_cast<const float4*>(s_img + c * PL + so);
    float4 u[C];
#pragma unroll
    for (int c = 0; c < C; ++c) u[c] = *reinterpret_cast<const float4*>(s_val + c * PL + so);
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = make_float4(v[c].x * sc, v[c].y * sc, v[c].z * sc, v[c].w * sc);
    if (!row_in || xb < 0 || xb + 3 >= W) {  // some element lies outside the image
      if (!row_in || xb + 0 < 0 || xb + 0 >= W) v[0].x = PP_SENTINEL;
      if (!row_in || xb + 1 < 0 || xb + 1 >= W) v[0].y = PP_SENTINEL;
      if (!row_in || xb + 2 < 0 || xb + 2 >= W) v[0].z = PP_SENTINEL;
      if (!row_in || xb + 3 < 0 || xb + 3 >= W) v[0].w = PP_SENTINEL;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(s_img + c * PL + so) = v[c];
    float w[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) w[c][0] = u[c].x, w[c][1] = u[c].y, w[c][2] = u[c].z, w[c][3] = u[c].w;
    if (CS != C) {  // p0 = 1 / (1 + e^(v1 - v0))
#pragma unroll
      for (int e = 0; e < 4; ++e) w[0][e] = rcp_approx(1.f + ex2_approx((w[1][e] - w[0][e]) * LOG2E));
    } else if (SOFTMAX) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float m = w[0][e];
#pragma unroll
        for (int c = 1; c < C; ++c) m = fmaxf(m, w[c][e]);
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          w[c][e] = ex2_approx((w[c][e] - m) * LOG2E);
          s += w[c][e];
        }
        const float inv = rcp_approx(s);
#pragma unroll
        for (int c = 0; c < C; ++c) w[c][e] *= inv;
      }
    }
    if (CS != C || SOFTMAX) {
#pragma unroll
      for (int c = 0; c < CS; ++c)
        *reinterpret_cast<float4*>(s_val + c * PL + so) = make_float4(w[c][0], w[c][1], w[c][2], w[c][3]);
    }
  }
}

template <int CS>
__device__ __forceinline__ void pp_zero(float (&X)[8][CS]) {
#pragma unroll
  for (int w = 0; w < 8; ++w)
#pragma unroll
    for (int c = 0; c < CS; ++c) X[w][c] = 0.f;
}


template <int C, bool SOFTMAX>
struct PpCfg {
  static constexpr int CS = (C == 2 && SOFTMAX) ? 1 : C;  // accumulated channels
  static constexpr int ROWS = PP_SEGS * 2 + 2;            // staged rows: 2 warm-up + owned + 2 look-ahead
  static constexpr int CAP = PP_SEGS * 2 - 2;             // owned rows per tile
  static constexpr int PLANE = (ROWS * PP_PITCH + 31) / 32 * 32;  // floats; planes start 128-byte aligned (TMA)
  static constexpr int BUF = (3 + C) * PLANE;                     // one tile buffer: image planes | value planes
  static constexpr int CARRY = (PP_SEGS - 1) * 2 * CS * 64;       // what a segment adds to the two rows of the next
  static constexpr int GB = (6 * CS * CAP + 31) / 32 * 32;        // G of the band-column pixels
  static constexpr size_t smem_bytes = (2 * (size_t)BUF + CARRY + GB + 64) * sizeof(float);
};

__device__ __forceinline__ void pp_bar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
          pp_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void pp_bar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pp_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pp_sync_m() { asm volatile("bar.sync 1, %0;" ::"n"(PP_M_THREADS) : "memory"); }
__device__ __forceinline__ void pp_sync_h() { asm volatile("bar.sync 2, %0;" ::"n"(PP_H_THREADS) : "memory"); }

__device__ __forceinline__ PpBlk pp_tile(const PpParams& Q, int tile) {
  PpBlk K;
  const int rb = tile % Q.nb, rest = tile / Q.nb;
  const int tx = rest % Q.n_x;
  K.b = rest / Q.n_x;
  K.x0 = tx * PP_TW;
  const int H = Q.p.H, base = H / Q.nb, extra = H - base * Q.nb;  // the first H % nb blocks have one row more
  K.ys = rb * base + min(rb, extra);
  K.n = base + (rb < extra ? 1 : 0);
  K.nc = K.n + 2;
  K.xband = (K.x0 <= 2) || (min(K.x0 + PP_TW, Q.p.W) - 1 >= Q.p.W - 3);
  K.scale2 = (float)(4.0 * Q.p.kappa) * (Q.p.grad_out ? __ldg(Q.p.grad_out + (Q.p.per_image ? K.b : 0)) : 1.f);
  return K;
}

// owned pixels / band-column pixels among the 4 columns of a strip
__device__ __forceinline__ void pp_masks(const PpParams& Q, const PpBlk& K, int strip, int& okmask, int& bandmask) {
  okmask = 0, bandmask = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 4 * strip + j;  // staged column: 0,1 and 62,63 are halo
    const int x = K.x0 - 2 + col;
    if (col >= 2 && col < 2 + PP_TW && x < Q.p.W) {
      okmask |= 1 << j;
      if (K.xband && pp_band_slot(x, Q.p.W) >= 0) bandmask |= 1 << j;
    }
  }
}

#ifdef WSDL_PS_TRACE  // debug aid (scripts/trace_pipe.py): timestamps per CTA, role, tile, phase
__device__ unsigned long long pp_trace_buf[512 * 2 * 8 * 4];
#define PP_TR(role_, j_, slot_)                                                                  \
  do {                                                                                           \
    if ((threadIdx.x == 0 || threadIdx.x == PP_M_THREADS) && blockIdx.x < 512 && (j_) < 8) {     \
      unsigned long long t__;                                                                    \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                    \
      pp_trace_buf[((blockIdx.x * 2 + (role_)) * 8 + (j_)) * 4 + (slot_)] = t__;                 \
    }                                                                                            \
  } while (0)
#else
#define PP_TR(role_, j_, slot_)
#endif

template <int C, bool SOFTMAX>
__global__ void __launch_bounds__(PP_THREADS, PP_CTAS_PER_SM)
    pairwise_pipe_kernel(const __grid_constant__ PpParams Q, const __grid_constant__ CUtensorMap tm_img,
                         const __grid_constant__ CUtensorMap tm_val) {
  using Cfg = PpCfg<C, SOFTMAX>;
  constexpr int CS = Cfg::CS, PL = Cfg::PLANE, CAP = Cfg::CAP;
  extern __shared__ __align__(128) float pp_smem[];
  float* s_carry = pp_smem + 2 * Cfg::BUF;  // [15][2][CS][64]
  float* s_gband = s_carry + Cfg::CARRY;    // [6][CS][CAP]
  float* s_wx = s_gband + Cfg::GB;          // [6][2][5]: column weights of the 6 band slots
  __shared__ __align__(8) unsigned long long s_full[2], s_ready[2], s_marched[2];
  __shared__ float s_red[2][PP_M_THREADS / 32];  // [tile parity][warp]
  __shared__ double s_dred[PP_M_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31;
  const int H = Q.p.H, W = Q.p.W;
  const int G = gridDim.x;
  const int n_my = (Q.n_tiles - (int)blockIdx.x + G - 1) / G;  // tiles blockIdx.x, blockIdx.x + G, ...
  const int kpi = Q.nb * Q.n_x;                                // tiles (= loss partials) per image

  if (tid == 0) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_full[b])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pp_smem_u32(&s_ready[b])), "n"(PP_H_THREADS / 32));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_marched[b])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 60) {  // column weights of the band slots: Wx(x -> x + j - 2), Wx(x + j - 2 -> x)
    const int slot = tid / 10, rem = tid - slot * 10, j = rem % 5;
    const int x = slot < 3 ? slot : W - 6 + slot, xb = x + j - 2;
    float w = 0.f;
    if (xb >= 0 && xb < W) w = rem < 5 ? pp_w1d(x, xb, W, Q.g1, Q.g4) : pp_w1d(xb, x, W, Q.g1, Q.g4);
    s_wx[tid] = w;
  }
  __syncthreads();  // the only CTA-wide barrier: from here on the two roles run on their own

  if (tid >= PP_M_THREADS) {
    // =============================== H: load and convert ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PP_H_REGS));
    const int gt = tid - PP_M_THREADS;
    auto issue = [&](int j) {  // leader only: TMA load of this CTA's j-th tile into buffer j & 1
      const int b = j & 1;
      const PpBlk K = pp_tile(Q, (int)blockIdx.x + j * G);
      float* planes = pp_smem + b * Cfg::BUF;
      const unsigned bar = pp_smem_u32(&s_full[b]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the planes were last written by threads
      const unsigned bytes = (unsigned)((3 + C) * Cfg::ROWS * PP_PITCH * sizeof(float));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
      for (int c = 0; c < 3 + C; ++c) {
        const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
        const int pl = c < 3 ? K.b * 3 + c : K.b * C + (c - 3);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                pp_smem_u32(planes + c * PL)),
            "l"(tm), "r"(K.x0 - 4), "r"(K.ys - 2), "r"(pl), "r"(bar)
            : "memory");
      }
    };
    if (Q.use_tma && gt == 0) {
      issue(0);
      if (n_my > 1) issue(1);
    }
    for (int j = 0; j < n_my; ++j) {
      const int b = j & 1;
      const PpBlk K = pp_tile(Q, (int)blockIdx.x + j * G);
      float* s_img = pp_smem + b * Cfg::BUF;
      float* s_p = s_img + 3 * PL;
      const int rows = K.nc + 2;
      if (j >= 2) {  // M has finished with the tile that was in this buffer
        pp_bar_wait(&s_marched[b], ((j - 2) >> 1) & 1);
        if (Q.use_tma && gt == 0) issue(j);
      }
      if (Q.use_tma) {
        pp_bar_wait(&s_full[b], (j >> 1) & 1);
      } else {
        pp_tile_load_slow<C, PL>(Q, K, s_img, s_p, rows, gt, PP_H_THREADS);
        pp_sync_h();
      }
      PP_TR(1, j, 0);
      pp_tile_transform<C, CS, SOFTMAX, PL>(Q, K, s_img, s_p, rows, gt, PP_H_THREADS);
      __syncwarp();
      if (lane == 0) pp_bar_arrive(&s_ready[b]);
      PP_TR(1, j, 1);
    }
    return;
  }

  // =============================== M: march ===============================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PP_M_REGS));
  const int gw = tid >> 5;
  const int seg = gw * 2 + (lane >> 4), strip = lane & 15;
  const float ksu = Q.p.ks_unit;
  for (int j = 0; j < n_my; ++j) {
    const int b = j & 1, tile = (int)blockIdx.x + j * G;
    const PpBlk K = pp_tile(Q, tile);
    float* s_img = pp_smem + b * Cfg::BUF;
    float* s_p = s_img + 3 * PL;
    int okmask, bandmask;
    pp_masks(Q, K, strip, okmask, bandmask);
    float lsum = 0.f;
    pp_bar_wait(&s_ready[b], (j >> 1) & 1);
    PP_TR(0, j, 0);

    // two centre rows per segment: t0, t0 + 1.  Their G stays in registers until the segment above has handed over
    // what it adds to them.
    const int t0 = seg * 2;
    float g0[4][CS], g1[4][CS], p0[4][CS], p1[4][CS];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < CS; ++c) g0[q][c] = 0.f, g1[q][c] = 0.f, p0[q][c] = 0.f, p1[q][c] = 0.f;
    if (gw * 4 < K.nc) {  // warp-uniform: at least one of its two segments has rows
      float A[8][CS], Bq[8][CS], Cq[8][CS];
      pp_zero<CS>(A), pp_zero<CS>(Bq), pp_zero<CS>(Cq);
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int t = t0 + s;
        if (t < K.nc) {
          // exponent offsets of this centre row: spatial term + multiplicity of the row pairs at the top / bottom border
          const int y = K.ys - 2 + t;
          const float l0 = (y == 1 || y == H - 2) ? Q.l1g : 0.f;
          const float l1 = (y == 0 || y == H - 2) ? Q.l32 : 0.f;
          const float l2 = (y == 0 || y == H - 3) ? Q.l32 : 0.f;
          PpKs ks;
          ks.a1 = ksu + l0, ks.a4 = 4.f * ksu + l0;
          ks.b0 = ksu + l1, ks.b1 = 2.f * ksu + l1, ks.b4 = 5.f * ksu + l1;
          ks.c0 = 4.f * ksu + l2, ks.c1 = 5.f * ksu + l2, ks.c4 = 8.f * ksu + l2;
          if (s == 0) pp_step<CS, PL>(A, Bq, Cq, p0, s_img, s_p, t * PP_PITCH + 4 * strip, ks);
          else pp_step<CS, PL>(Bq, Cq, A, p1, s_img, s_p, t * PP_PITCH + 4 * strip, ks);
        }
        if (s == 0) {
          pp_exchange<CS>(A, g0, strip);
          pp_zero<CS>(A);  // becomes the accumulator of row t0 + 3
        } else {
          pp_exchange<CS>(Bq, g1, strip);
        }
      }
      // rows t0 + 2 (Cq) and t0 + 3 (A) belong to the next segment: hand over what this one contributed
      float oy[4][CS], oz[4][CS];
      pp_exchange<CS>(Cq, oy, strip);
      pp_exchange<CS>(A, oz, strip);
      if (seg < PP_SEGS - 1) {
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          if (t0 + 2 < K.nc)
            *reinterpret_cast<float4*>(s_carry + ((seg * 2 + 0) * CS + c) * 64 + 4 * strip) =
                make_float4(oy[0][c], oy[1][c], oy[2][c], oy[3][c]);
          if (t0 + 3 < K.nc)
            *reinterpret_cast<float4*>(s_carry + ((seg * 2 + 1) * CS + c) * 64 + 4 * strip) =
                make_float4(oz[0][c], oz[1][c], oz[2][c], oz[3][c]);
        }
      }
    }
    PP_TR(0, j, 1);
    pp_sync_m();  // the carries are in shared memory
    if (seg > 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int t = t0 + r;
        if (t < K.nc) {
          float(&g)[4][CS] = r == 0 ? g0 : g1;
          const float(&pc)[4][CS] = r == 0 ? p0 : p1;
#pragma unroll
          for (int c = 0; c < CS; ++c) {
            const float4 k = *reinterpret_cast<const float4*>(s_carry + (((seg - 1) * 2 + r) * CS + c) * 64 + 4 * strip);
            g[0][c] += k.x, g[1][c] += k.y, g[2][c] += k.z, g[3][c] += k.w;
          }
          if (bandmask) {  // the band pass finishes these pixels from their G
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if ((bandmask >> q) & 1) {
                const int slot = pp_band_slot(K.x0 - 2 + 4 * strip + q, W);
#pragma unroll
                for (int c = 0; c < CS; ++c) s_gband[(slot * CS + c) * CAP + (t - 2)] = g[q][c];
              }
          }
          pp_emit<C, CS, SOFTMAX>(Q, K, t, strip, okmask, bandmask, g, pc, lsum);
        }
      }
    }
    if (K.xband) {
      // ---- band columns (and corners): G + weight correction -> the gradient of these pixels ----
      pp_sync_m();  // band G is complete
      const int xe = min(K.x0 + PP_TW, W);
      const int nlo = max(0, min(3, xe) - K.x0), hi0 = max(W - 3, K.x0), nhi = max(0, xe - hi0);
      const int ncb = nlo + nhi;
      const size_t plane = (size_t)H * W;
      for (int i = tid; i < ncb * K.n; i += PP_M_THREADS) {  // column fastest
        const int ty = i / ncb, k = i - ty * ncb;
        const int x = k < nlo ? K.x0 + k : hi0 + (k - nlo), y = K.ys + ty;
        const int slot = pp_band_slot(x, W);
        float acc[CS], pz[CS], o[C];
        pp_xfix_item<CS, PL>(Q, s_img, s_p, s_wx + slot * 10, K.ys, K.x0, y, x, acc, pz);
#pragma unroll
        for (int c = 0; c < CS; ++c) acc[c] += s_gband[(slot * CS + c) * CAP + ty];
        lsum += pp_pixel_grad<C, CS, SOFTMAX>(K.scale2, pz, acc, o);
        if (Q.p.grad_values) {
          float* go = Q.p.grad_values + (size_t)K.b * C * plane + (size_t)y * W + x;
#pragma unroll
          for (int c = 0; c < C; ++c) go[c * plane] = o[c];
        }
      }
    }
    PP_TR(0, j, 2);
    const float w = warp_sum(lsum);
    if (lane == 0) s_red[j & 1][gw] = w;
    pp_sync_m();  // nobody reads the planes, the carries or the band G of this tile any more
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < PP_M_THREADS / 32; ++i) t += s_red[j & 1][i];
      const int img = tile / kpi, within = tile - img * kpi;
      __stcg(Q.p.partial + (size_t)img * kpi + within, 2.f * t);
      pp_bar_arrive(&s_marched[b]);
    }
    PP_TR(0, j, 3);
  }

  // ---- ticket: every CTA checks in; the M warps of the last CTA add the partials per image, in double ----
  if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(Q.p.ticket) : "memory");
  if ((int)blockIdx.x != G - 1) return;
  if (tid == 0) {
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(Q.p.ticket) : "memory");
      if (seen < (unsigned)G) __nanosleep(200);
    } while (seen < (unsigned)G);
    *Q.p.ticket = 0u;  // every CTA has checked in: leave the ticket ready for the next launch
  }
  pp_sync_m();
  double wtot = 0.0;
  for (int b = gw; b < Q.p.B; b += PP_M_THREADS / 32) {
    double acc = 0.0;
    for (int i = lane; i < kpi; i += 32) acc += (double)ld_cg_f32(Q.p.partial + (size_t)b * kpi + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (Q.p.per_image) {
      if (lane == 0) Q.p.loss_out[b] = (float)(acc * Q.p.kappa);
    } else {
      wtot += acc;
    }
  }
  if (!Q.p.per_image) {
    if (lane == 0) s_dred[gw] = wtot;
    pp_sync_m();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < PP_M_THREADS / 32; ++i) t += s_dred[i];
      Q.p.loss_out[0] = (float)(t * Q.p.kappa);
    }
  }
}

// Row blocks per column tile: at most `cap` rows each; among the next few candidates the count with the lowest
// modelled time = (march steps + fixed cost of a tile, in steps) x tiles per persistent CTA.
static int pp_row_blocks(int B, int H, int W, int cap) {
  const int n_x = (W + PP_TW - 1) / PP_TW;
  const int nb_min = (H + cap - 1) / cap;
  const double slots = (double)WSDL_NUM_SMS * PP_CTAS_PER_SM;
  int best = nb_min;
  double best_cost = 1e300;
  for (int nb = nb_min; nb <= nb_min + 12; ++nb) {
    const int n = (H + nb - 1) / nb;  // rows of the largest block
    if (n < 6 && nb > nb_min) break;
    const int S = 2;  // two rows per segment, always
    const double per_cta = (double)B * n_x * nb / slots;
    const double rounds = per_cta < 1.0 ? 1.0 : (double)(long long)(per_cta + 0.999999);
    const double cost = (S + 1.0) * rounds + 1.5;  // + the exposed load and conversion of the first tile
    if (cost < best_cost - 1e-9) best_cost = cost, best = nb;
  }
  return best;
}

static int pp_cap(int, int) { return PP_SEGS * 2 - 2; }

typedef CUresult (*PpEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PpEncodeFn pp_encoder() {  // cuTensorMapEncodeTiled through the runtime: no link-time dependency on libcuda
  static const PpEncodeFn fn = []() -> PpEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (PpEncodeFn)p;
  }();
  return fn;
}

// (W, H, planes) f32 tensor, box = 68 columns x `rows` rows x 1 plane, zero fill outside
static bool pp_encode(CUtensorMap* tm, const float* base, int W, int H, long long planes, int rows) {
  const PpEncodeFn enc = pp_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)PP_PITCH, (cuuint32_t)rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, bool SOFTMAX>
static int pp_launch_t(PpParams& Q, const CUtensorMap& tm_img, const CUtensorMap& tm_val, cudaStream_t s) {
  constexpr size_t smem = PpCfg<C, SOFTMAX>::smem_bytes;
  // function attributes are per device: remember which devices have them (idempotent; a race only repeats the call)
  static bool attr_done[64] = {};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, attr_done[0] = false;
  bool& attr_set = attr_done[dev_id];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_pipe_kernel<C, SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(pairwise_pipe_kernel<C, SOFTMAX>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int slots = WSDL_NUM_SMS * PP_CTAS_PER_SM;
  const int grid = Q.n_tiles < slots ? Q.n_tiles : slots;
  pairwise_pipe_kernel<C, SOFTMAX><<<grid, PP_THREADS, smem, s>>>(Q, tm_img, tm_val);
  WSDL_LAUNCH_CHECK();
  return 0;
}

}  // namespace pp

size_t pp_workspace_floats(int B, int H, int W) {
  const int n_x = (W + pp::PP_TW - 1) / pp::PP_TW;
  return (size_t)B * n_x * pp::pp_row_blocks(B, H, W, pp::PP_SEGS * 2 - 2);
}

// returns 1 when the shape is not this kernel's (the caller then takes another one), 0 on success
int pp_launch(const PwParams& P, cudaStream_t s) {
  using namespace pp;
  if (P.pad != 2 || P.C < 1 || P.C > 2 || P.H < 6 || P.W < 6) return 1;
  PpParams Q;
  Q.p = P;
  Q.n_x = (P.W + PP_TW - 1) / PP_TW;
  Q.nb = pp_row_blocks(P.B, P.H, P.W, pp_cap(P.C, P.inner_softmax));
  const long long n_tiles = (long long)Q.nb * Q.n_x * P.B;
  if (n_tiles > 0x3fffffffLL) return 1;
  Q.n_tiles = (int)n_tiles;
  Q.S = 2;
  Q.vec2_ok = ((P.W & 1) == 0) && (!P.grad_values || ((uintptr_t)P.grad_values & 7) == 0);
  Q.img_scale = sqrtf(-P.kc);
  Q.g1 = expf(-P.inv_2ss), Q.g4 = expf(-4.f * P.inv_2ss);
  Q.l32 = log2f(1.5f), Q.l1g = log2f(1.f + Q.g4);
  CUtensorMap tm_img, tm_val;
  memset(&tm_img, 0, sizeof(tm_img)), memset(&tm_val, 0, sizeof(tm_val));
  static const int no_tma = []() { const char* e = getenv("WSDL_PAIRWISE_NO_TMA"); return (e && e[0] == '1') ? 1 : 0; }();
  Q.use_tma = !no_tma && ((P.W & 3) == 0) && (((uintptr_t)P.values & 15) == 0) && (((uintptr_t)P.images & 15) == 0) &&
              pp_encode(&tm_img, P.images, P.W, P.H, 3LL * P.B, PP_SEGS * 2 + 2) &&
              pp_encode(&tm_val, P.values, P.W, P.H, (long long)P.C * P.B, PP_SEGS * 2 + 2);
  if (P.C == 2)
    return P.inner_softmax ? pp_launch_t<2, true>(Q, tm_img, tm_val, s) : pp_launch_t<2, false>(Q, tm_img, tm_val, s);
  return P.inner_softmax ? pp_launch_t<1, true>(Q, tm_img, tm_val, s) : pp_launch_t<1, false>(Q, tm_img, tm_val, s);
}

}  // namespace wsdl

#ifdef WSDL_PS_TRACE
extern "C" int wsdl_pp_trace_read(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, wsdl::pp::pp_trace_buf, (size_t)n * sizeof(unsigned long long));
}
#endif
