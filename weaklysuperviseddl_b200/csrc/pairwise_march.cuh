// Device code shared by the pair-symmetric kernels (pairwise_sym.cu: one tile per CTA; pairwise_stream.cu:
// persistent CTAs that stream tiles): tile geometry, the packed march step, exchange, border weights.
// See pairwise_sym.cu for the derivations (pair symmetry, Euler loss, border multiplicities, sentinel colour).
#pragma once
#include <cuda.h>
#include <string.h>

#include "pairwise.cuh"

namespace wsdl {

constexpr int PS_TW = 60;                      // owned columns per tile
constexpr int PS_PITCH = 68;                   // smem row: image x0-4 .. x0+63 (2 pad + 2 halo | 60 | 2 halo + 2 pad)
constexpr int PS_Q = PS_PITCH / 4;             // float4 per staged row
#ifndef WSDL_PS_SEGS
#define WSDL_PS_SEGS 8
#endif
constexpr int PS_SEGS = WSDL_PS_SEGS;          // row segments per block, one per half warp (even)
constexpr int PS_THREADS = 16 * PS_SEGS;
constexpr int PS_WARPS = PS_SEGS / 2;
// Tile height and residency of the single-loss kernels and the tile-streaming kernel.  3 CTAs per SM with segments of
// up to 5 rows (40 centre rows, 67 KB of shared memory, <= 168 registers) beat 4 CTAs with 4-row segments (<= 128
// registers) on configs[1] -- the fixed cost of a block (tile load, conversion, segment heads, loss partial) is spread
// over more rows, 224 rows split into 6 blocks of 38 use 93 % of the centre rows instead of 87.5 %, and the packed march
// needs no spills.  (The fused kernel has 6-row segments: its own constants, DU_*, in pairwise_sym.cu.)
#ifndef WSDL_PS_SMAX
#define WSDL_PS_SMAX 5
#endif
#ifndef WSDL_PS_CTAS
#define WSDL_PS_CTAS 3
#endif
#ifndef WSDL_PS_PACKED
#define WSDL_PS_PACKED 1  // two-lane FP32 instructions (FADD2 / FFMA2) in the march of every variant
#endif
constexpr int PS_SMAX = WSDL_PS_SMAX;           // rows per segment
constexpr int PS_CENTERS = PS_SEGS * PS_SMAX;  // 40 centre rows per block: 2 warm-up + 38 owned
constexpr int PS_ROWS = PS_CENTERS + 2;        // + 2 look-ahead rows
constexpr int PS_CAP = PS_CENTERS - 2;         // owned rows per block
constexpr int PS_PLANE = (PS_ROWS * PS_PITCH + 31) / 32 * 32;  // floats; planes start 128-byte aligned (TMA)

struct PsParams {
  PwParams p;
  int n_x;          // column tiles per image
  int nb;           // row blocks per column tile; grid = (nb, n_x, B)
  int S;            // rows per segment of this launch: ceil((rows of the largest block + 2) / 8), >= 2
  int use_tma;      // W % 4 == 0, 16-byte aligned inputs, tensor maps encoded
  unsigned stagger_ns;  // hold-back of the first wave's tile loads, per CTA already resident on the SM
  int vec2_ok;      // W % 2 == 0 and 8-byte aligned gradient: float2 stores
  float img_scale;  // sqrt(-kc)
  float g1, g4;     // gamma, gamma^4 (gamma = exp(-1 / (2 sigma_space^2)), 1 without a spatial term)
  float l32, l1g;   // log2(3/2), log2(1 + gamma^4): row-border pair multiplicities as exponent offsets
  unsigned* slots;  // single-loss kernels: one loss partial per CTA (a float's bits), 0xffffffff when empty
  float kappa4;     // (float)(4 kappa): rounded on the host (no double-precision arithmetic in the CTAs' prologue)
  // single-loss kernels: weight tables of the band-column pass (ps_band_tables1, H >= 10), as PsDual's but for one loss.
  // fx1 [6 slots][2][4]: partner-column offset, 1/2 Wx(a->b), 1/2 Wx(b->a), 1/2 gamma^dx^2; wy1 [11][5][4]: forward,
  // backward and applied-by-the-march row weights of the row-band rows 0..4, H-5..H-1 and of any other row.
  float fx1[6 * 2 * 4];
  float wy1[11 * 5 * 4];
};

constexpr float PS_SENTINEL = 1e15f;  // staged colour (channel 0) of positions outside the image: 2^-(1e30) = 0

struct PsKs {  // exponent offsets of one row step: spatial term + row-border multiplicity
  float a1, a4;      // partner in the same row, dx^2 = 1, 4
  float b0, b1, b4;  // one row down,  dx^2 = 0, 1, 4
  float c0, c1, c4;  // two rows down
};

template <int C>
struct PsWin {  // an 8-column window of one staged row: columns 4*strip .. 4*strip+7 of the smem row
  float i[3][8];
  float p[C][8];
};

template <int C>
__device__ __forceinline__ void ps_load(PsWin<C>& w, const float* s_img, const float* s_p, int off) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(s_img + c * PS_PLANE + off);
    const float4 b = *reinterpret_cast<const float4*>(s_img + c * PS_PLANE + off + 4);
    w.i[c][0] = a.x, w.i[c][1] = a.y, w.i[c][2] = a.z, w.i[c][3] = a.w;
    w.i[c][4] = b.x, w.i[c][5] = b.y, w.i[c][6] = b.z, w.i[c][7] = b.w;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(s_p + c * PS_PLANE + off);
    const float4 b = *reinterpret_cast<const float4*>(s_p + c * PS_PLANE + off + 4);
    w.p[c][0] = a.x, w.p[c][1] = a.y, w.p[c][2] = a.z, w.p[c][3] = a.w;
    w.p[c][4] = b.x, w.p[c][5] = b.y, w.p[c][6] = b.z, w.p[c][7] = b.w;
  }
}

// one unordered pair: k = 2^(ks - |I'(a) - I'(b)|^2);  G(a) += k (p(a) - p(b));  G(b) -= k (p(a) - p(b))
template <int C>
__device__ __forceinline__ void ps_pair(float (&ga)[C], float (&gb)[C], const PsWin<C>& a, int ia, const PsWin<C>& b,
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
template <int C>
__device__ __forceinline__ void ps_step(float (&X)[8][C], float (&Y)[8][C], float (&Z)[8][C], float (&pc)[4][C],
                                        const float* s_img, const float* s_p, int off, const PsKs& ks) {
  PsWin<C> c;
  ps_load<C>(c, s_img, s_p, off);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int cc = 0; cc < C; ++cc) pc[j][cc] = c.p[cc][2 + j];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    ps_pair<C>(X[2 + j], X[3 + j], c, 2 + j, c, 3 + j, ks.a1);
    ps_pair<C>(X[2 + j], X[4 + j], c, 2 + j, c, 4 + j, ks.a4);
  }
  PsWin<C> n;
  ps_load<C>(n, s_img, s_p, off + PS_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx)
      ps_pair<C>(X[2 + j], Y[2 + j + dx], c, 2 + j, n, 2 + j + dx, dx == 0 ? ks.b0 : (dx * dx == 1 ? ks.b1 : ks.b4));
  ps_load<C>(n, s_img, s_p, off + 2 * PS_PITCH);
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx)
      ps_pair<C>(X[2 + j], Z[2 + j + dx], c, 2 + j, n, 2 + j + dx, dx == 0 ? ks.c0 : (dx * dx == 1 ? ks.c1 : ks.c4));
}

// A finished accumulator row: the two columns either side of the strip belong to the neighbouring lanes.
template <int C>
__device__ __forceinline__ void ps_exchange(const float (&X)[8][C], float (&own)[4][C], int strip) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float r0 = __shfl_down_sync(0xffffffffu, X[0][c], 1), r1 = __shfl_down_sync(0xffffffffu, X[1][c], 1);
    const float l0 = __shfl_up_sync(0xffffffffu, X[6][c], 1), l1 = __shfl_up_sync(0xffffffffu, X[7][c], 1);
    // strip 0 receives nothing real from its left and strip 15 nothing from its right (the lane there belongs to the
    // other segment): the two sums that take those values are the tile's halo columns, which no caller keeps
    own[0][c] = X[2][c] + l0;
    own[1][c] = X[3][c] + l1;
    own[2][c] = X[4][c] + r0;
    own[3][c] = X[5][c] + r1;
  }
}

// ---- packed (f32x2) form of the step: see the dual kernel below for the layout (even / odd column pairs) ----
__device__ __forceinline__ float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }

template <int CS>
struct PsWinP {
  float2 e[3 + CS][4];  // [plane: I0, I1, I2, p...][even pair = columns (2k, 2k+1) of the 8-column window]
};

template <int CS, int PLANE = PS_PLANE>
__device__ __forceinline__ void ps_loadp(PsWinP<CS>& w, const float* s_img, const float* s_p, int off) {
#pragma unroll
  for (int c = 0; c < 3 + CS; ++c) {
    const float* src = (c < 3 ? s_img + c * PLANE : s_p + (c - 3) * PLANE) + off;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 b = *reinterpret_cast<const float4*>(src + 4);
    w.e[c][0] = make_float2(a.x, a.y), w.e[c][1] = make_float2(a.z, a.w);
    w.e[c][2] = make_float2(b.x, b.y), w.e[c][3] = make_float2(b.z, b.w);
  }
}

// two pairs with the same offset: k = 2^(ks - |I'(a) - I'(b)|^2);  G(a) += k (p(a) - p(b));  G(b) -= k (p(a) - p(b))
template <int CS>
__device__ __forceinline__ void ps_pairp(float2 (&ga)[CS], float2 (&gb)[CS], const float2 (&a)[3 + CS],
                                         const float2 (&b)[3 + CS], float ks) {
  const float2 d0 = __fadd2_rn(a[0], f2neg(b[0])), d1 = __fadd2_rn(a[1], f2neg(b[1])), d2 = __fadd2_rn(a[2], f2neg(b[2]));
  const float2 e = __ffma2_rn(f2neg(d2), d2, __ffma2_rn(f2neg(d1), d1, __ffma2_rn(f2neg(d0), d0, make_float2(ks, ks))));
  const float2 k = make_float2(ex2_approx(e.x), ex2_approx(e.y));
#pragma unroll
  for (int c = 0; c < CS; ++c) {
    const float2 dp = __fadd2_rn(a[3 + c], f2neg(b[3 + c]));
    ga[c] = __ffma2_rn(k, dp, ga[c]);
    gb[c] = __ffma2_rn(f2neg(k), dp, gb[c]);
  }
}

// All 12 forward pairs of the columns of row t; X, Y, Z: even-pair accumulators [pair][channel] of rows t, t+1, t+2.
template <int CS>
__device__ __forceinline__ void ps_stepp(float2 (&X)[4][CS], float2 (&Y)[4][CS], float2 (&Z)[4][CS], float (&pc)[4][CS],
                                         const float* s_img, const float* s_p, int off, const PsKs& ks) {
  PsWinP<CS> c;
  ps_loadp<CS>(c, s_img, s_p, off);
#pragma unroll
  for (int cc = 0; cc < CS; ++cc)
    pc[0][cc] = c.e[3 + cc][1].x, pc[1][cc] = c.e[3 + cc][1].y, pc[2][cc] = c.e[3 + cc][2].x, pc[3][cc] = c.e[3 + cc][2].y;
  float2 co[3][3 + CS], ce[4][3 + CS];  // odd pairs (2k+1, 2k+2) and even pairs of row t, by plane
#pragma unroll
  for (int pl = 0; pl < 3 + CS; ++pl) {
#pragma unroll
    for (int k = 0; k < 3; ++k) co[k][pl] = make_float2(c.e[pl][k].y, c.e[pl][k + 1].x);
#pragma unroll
    for (int k = 0; k < 4; ++k) ce[k][pl] = c.e[pl][k];
  }
  float2 XO[3][CS];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int cc = 0; cc < CS; ++cc) XO[k][cc] = make_float2(0.f, 0.f);
  ps_pairp<CS>(X[1], XO[1], ce[1], co[1], ks.a1);  // same row, dx = 1
  ps_pairp<CS>(X[2], XO[2], ce[2], co[2], ks.a1);
  ps_pairp<CS>(X[1], X[2], ce[1], ce[2], ks.a4);   // dx = 2
  ps_pairp<CS>(X[2], X[3], ce[2], ce[3], ks.a4);
#pragma unroll
  for (int r = 1; r <= 2; ++r) {
    float2 (&Yr)[4][CS] = r == 1 ? Y : Z;
    const float k0 = r == 1 ? ks.b0 : ks.c0, k1 = r == 1 ? ks.b1 : ks.c1, k4 = r == 1 ? ks.b4 : ks.c4;
    PsWinP<CS> n;
    ps_loadp<CS>(n, s_img, s_p, off + r * PS_PITCH);
    float2 ne[4][3 + CS];
#pragma unroll
    for (int pl = 0; pl < 3 + CS; ++pl)
#pragma unroll
      for (int k = 0; k < 4; ++k) ne[k][pl] = n.e[pl][k];
    ps_pairp<CS>(X[1], Yr[0], ce[1], ne[0], k4);   // dx = -2
    ps_pairp<CS>(X[2], Yr[1], ce[2], ne[1], k4);
    ps_pairp<CS>(XO[1], Yr[1], co[1], ne[1], k1);  // dx = -1: centres 3,4 and 5,6 (grouped by partner column)
    ps_pairp<CS>(XO[2], Yr[2], co[2], ne[2], k1);
    ps_pairp<CS>(X[1], Yr[1], ce[1], ne[1], k0);   // dx = 0
    ps_pairp<CS>(X[2], Yr[2], ce[2], ne[2], k0);
    ps_pairp<CS>(XO[0], Yr[1], co[0], ne[1], k1);  // dx = +1: centres 1,2 and 3,4
    ps_pairp<CS>(XO[1], Yr[2], co[1], ne[2], k1);
    ps_pairp<CS>(X[1], Yr[2], ce[1], ne[2], k4);   // dx = +2
    ps_pairp<CS>(X[2], Yr[3], ce[2], ne[3], k4);
  }
#pragma unroll
  for (int ch = 0; ch < CS; ++ch) {
    X[0][ch].y += XO[0][ch].x;
    X[1][ch].x += XO[0][ch].y;
    X[1][ch].y += XO[1][ch].x;
    X[2][ch].x += XO[1][ch].y;
    X[2][ch].y += XO[2][ch].x;
    X[3][ch].x += XO[2][ch].y;
  }
}

template <int CS>
__device__ __forceinline__ void ps_exchangep(const float2 (&X)[4][CS], float (&own)[4][CS], int strip) {
#pragma unroll
  for (int c = 0; c < CS; ++c) {
    const float r0 = __shfl_down_sync(0xffffffffu, X[0][c].x, 1), r1 = __shfl_down_sync(0xffffffffu, X[0][c].y, 1);
    const float l0 = __shfl_up_sync(0xffffffffu, X[3][c].x, 1), l1 = __shfl_up_sync(0xffffffffu, X[3][c].y, 1);
    // (strip 0 / strip 15: what arrives from the other segment's lane only reaches halo columns -- see ps_exchange)
    own[0][c] = X[1][c].x + l0;
    own[1][c] = X[1][c].y + l1;
    own[2][c] = X[2][c].x + r0;
    own[3][c] = X[2][c].y + r1;
  }
}

struct PsBlk {
  int b, x0, ys, n, nc;
  bool xband;    // the tile owns pixels within 3 columns of the left / right image border
  float scale2;  // 4 kappa * upstream gradient: dL/dp = scale2 * G
};

// band slot of a coordinate: 0..2 for the low band, 3..5 for the high band, -1 outside (needs n >= 6)
__device__ __forceinline__ int ps_band_slot(int v, int n) { return v <= 2 ? v : (v >= n - 3 ? v - (n - 6) : -1); }

// One pixel: dL/dvalue from G (linear in G: softmax backward included), and its term of 2 kappa sum (p - 1/2) G.
template <int C, int CS, bool SOFTMAX>
__device__ __forceinline__ float ps_pixel_grad(float scale2, const float (&p)[CS], const float (&g)[CS], float (&out)[C]) {
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

// Gradient of 4 finished pixels of centre row t: store, and accumulate the loss term.
template <int C, int CS, bool SOFTMAX>
__device__ __forceinline__ void ps_emit(const PsParams& Q, const PsBlk& K, int t, int strip, int okmask,
                                        const float (&G)[4][CS], const float (&pc)[4][CS], float* s_gband, float& lsum) {
  const int H = Q.p.H, W = Q.p.W;
  const int y = K.ys - 2 + t;
  const int xs = K.x0 - 2 + 4 * strip;  // image column of j = 0
  if (K.xband) {  // block-uniform: the band-column pass finishes these pixels from their G
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int slot = ps_band_slot(xs + j, W);
      if (((okmask >> j) & 1) && slot >= 0) {
#pragma unroll
        for (int c = 0; c < CS; ++c) s_gband[(slot * CS + c) * PS_CAP + (t - 2)] = G[j][c];
      }
    }
  }
  float out[C][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float o[C];
    const float l = ps_pixel_grad<C, CS, SOFTMAX>(K.scale2, pc[j], G[j], o);
#pragma unroll
    for (int c = 0; c < C; ++c) out[c][j] = o[c];
    if ((okmask >> j) & 1) lsum += l;
  }
  if (Q.p.grad_values) {
    const size_t plane = (size_t)H * W;
    float* go = Q.p.grad_values + (size_t)K.b * C * plane + (size_t)y * W + xs;
    if (Q.vec2_ok) {  // xs is even, so (xs, xs+1) and (xs+2, xs+3) are inside or outside W together
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
          if ((okmask >> j) & 1) go[c * plane + j] = out[c][j];
    }
  }
}

// gamma^(k^2) for k = 0, 1, 2
__host__ __device__ __forceinline__ float ps_gpow(int k, float g1, float g4) { return k == 0 ? 1.f : (k == 1 ? g1 : g4); }

// W(u -> v) = sum_{d=-2..2} [reflect(u + d) == v] gamma^(d^2) for u inside [0, n), n >= 6; 0 for v outside (closed
// form of pairwise.cu's axis_multiplicity for pad 2).
__host__ __device__ __forceinline__ float ps_w1d(int u, int v, int n, float g1, float g4) {
  if (v < 0 || v >= n) return 0.f;
  const int d = v > u ? v - u : u - v, s = u + v, e = 2 * (n - 1) - s;
  float w = d <= 2 ? ps_gpow(d, g1, g4) : 0.f;
  if (v >= 1 && s <= 2) w += ps_gpow(s, g1, g4);      // offset -s lands on -v, which reflects to v
  if (v <= n - 2 && e <= 2) w += ps_gpow(e, g1, g4);  // offset +e lands on 2(n-1) - v
  return w;
}

// multiplicity the march applies to a pair of rows (relative to the interior 2 gamma^|d|^2): see the row step
__host__ __device__ __forceinline__ float ps_row_mult(int ya, int yb, int H, float g4) {
  const int lo = ya < yb ? ya : yb, d = ya < yb ? yb - ya : ya - yb;
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
// escale: the staged image is scaled for one sigma_color; another loss on the same tile multiplies the squared
// distance by (sigma / sigma')^2.
template <int CS>
__device__ __forceinline__ void ps_xfix_item(int H, float g1, float g4, float escale, const float* s_img, const float* s_p,
                                            const float* s_wx, int ys, int x0, int zy, int zx, float (&acc)[CS],
                                            float (&pz)[CS]) {
  const int so = (zy - (ys - 2)) * PS_PITCH + (zx - (x0 - 4));
  const float i0 = s_img[so], i1 = s_img[PS_PLANE + so], i2 = s_img[2 * PS_PLANE + so];
#pragma unroll
  for (int c = 0; c < CS; ++c) pz[c] = s_p[c * PS_PLANE + so], acc[c] = 0.f;
  // the (at most two) partner columns whose weights are not the interior ones
  int jsp[2] = {-1, -1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float f = s_wx[j], b = s_wx[5 + j], g = ps_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
    if (f != 0.f && (f != g || b != g)) {
      if (jsp[0] < 0) jsp[0] = j;
      else jsp[1] = j;
    }
  }
  // row weights: forward, backward, and what the march applies (2 gamma^dy^2 times its row multiplicity)
  float wyf[5], wyb[5], wym[5];
  if (zy >= 5 && zy <= H - 6) {  // every partner row is clear of the row bands
#pragma unroll
    for (int i = 0; i < 5; ++i) wyf[i] = wyb[i] = ps_gpow(i < 2 ? 2 - i : i - 2, g1, g4), wym[i] = 2.f * wyf[i];
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yb = zy + i - 2;
      const bool in = yb >= 0 && yb < H;
      wyf[i] = in ? ps_w1d(zy, yb, H, g1, g4) : 0.f;
      wyb[i] = in ? ps_w1d(yb, zy, H, g1, g4) : 0.f;
      wym[i] = in ? 2.f * ps_gpow(i < 2 ? 2 - i : i - 2, g1, g4) * ps_row_mult(zy, yb, H, g4) : 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = jsp[u] < 0 ? 2 : jsp[u];  // no such column: the pixel's own (weight 0 below)
    const float live = jsp[u] < 0 ? 0.f : 0.5f;
    const float wxf = s_wx[j], wxb = s_wx[5 + j], gx = ps_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float diff = live * (fmaf(wyf[i], wxf, wyb[i] * wxb) - wym[i] * gx);  // 0 for rows outside the image
      const int sn = so + (i - 2) * PS_PITCH + (j - 2);
      const float d0 = i0 - s_img[sn], d1 = i1 - s_img[PS_PLANE + sn], d2 = i2 - s_img[2 * PS_PLANE + sn];
      const float k = diff * ex2_approx(escale * fmaf(-d2, d2, fmaf(-d1, d1, -d0 * d0)));
#pragma unroll
      for (int c = 0; c < CS; ++c) acc[c] = fmaf(k, pz[c] - s_p[c * PS_PLANE + sn], acc[c]);
    }
  }
}

// The same for the cut loss (gamma = 1, squared distances as staged) and the boundary loss (gamma_b, distances times
// `ratio`) of one band pixel at once: the two corrections differ in their weights only, so the partner loads, the colour
// differences and p(a) - p(b) are shared.  s_wxc / s_wxb: the pixel's column-weight rows of the two tables.
template <int PLANE = PS_PLANE>
__device__ __forceinline__ void ps_xfix_dual(int H, float g1b, float g4b, float ratio, const float* s_img, const float* s_p,
                                             const float* s_wxc, const float* s_wxb, int ys, int x0, int zy, int zx,
                                             float& ac, float& ab, float& pz) {
  const int so = (zy - (ys - 2)) * PS_PITCH + (zx - (x0 - 4));
  const float i0 = s_img[so], i1 = s_img[PLANE + so], i2 = s_img[2 * PLANE + so];
  pz = s_p[so];
  // the (at most two) partner columns whose weights are not the interior ones: reflect geometry, the same for both losses
  int jsp[2] = {-1, -1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const float f = s_wxc[j], b = s_wxc[5 + j];
    if (f != 0.f && (f != 1.f || b != 1.f)) {
      if (jsp[0] < 0) jsp[0] = j;
      else jsp[1] = j;
    }
  }
  float wc[5][3], wb[5][3];  // per partner row: forward, backward, applied by the march
  if (zy >= 5 && zy <= H - 6) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float g = ps_gpow(i < 2 ? 2 - i : i - 2, g1b, g4b);
      wc[i][0] = wc[i][1] = 1.f, wc[i][2] = 2.f;
      wb[i][0] = wb[i][1] = g, wb[i][2] = 2.f * g;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yb = zy + i - 2;
      const bool in = yb >= 0 && yb < H;
      const int dd = i < 2 ? 2 - i : i - 2;
      wc[i][0] = in ? ps_w1d(zy, yb, H, 1.f, 1.f) : 0.f;
      wc[i][1] = in ? ps_w1d(yb, zy, H, 1.f, 1.f) : 0.f;
      wc[i][2] = in ? 2.f * ps_row_mult(zy, yb, H, 1.f) : 0.f;
      wb[i][0] = in ? ps_w1d(zy, yb, H, g1b, g4b) : 0.f;
      wb[i][1] = in ? ps_w1d(yb, zy, H, g1b, g4b) : 0.f;
      wb[i][2] = in ? 2.f * ps_gpow(dd, g1b, g4b) * ps_row_mult(zy, yb, H, g4b) : 0.f;
    }
  }
  ac = 0.f, ab = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = jsp[u] < 0 ? 2 : jsp[u];  // no such column: the pixel's own (weight 0 below)
    const float live = jsp[u] < 0 ? 0.f : 0.5f;
    const float cf = s_wxc[j], cb = s_wxc[5 + j];
    const float bf = s_wxb[j], bb = s_wxb[5 + j], gx = ps_gpow(j < 2 ? 2 - j : j - 2, g1b, g4b);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float dc = live * (fmaf(wc[i][0], cf, wc[i][1] * cb) - wc[i][2]);  // 0 for rows outside the image
      const float db = live * (fmaf(wb[i][0], bf, wb[i][1] * bb) - wb[i][2] * gx);
      const int sn = so + (i - 2) * PS_PITCH + (j - 2);
      const float d0 = i0 - s_img[sn], d1 = i1 - s_img[PLANE + sn], d2 = i2 - s_img[2 * PLANE + sn];
      const float e = fmaf(-d2, d2, fmaf(-d1, d1, -d0 * d0));
      const float dp = pz - s_p[sn];
      ac = fmaf(dc * ex2_approx(e), dp, ac);
      ab = fmaf(db * ex2_approx(ratio * e), dp, ab);
    }
  }
}

// ---- staging ----
__device__ __forceinline__ unsigned ps_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// Rows [r0, r1) of the tile, element by element, raw values (0 outside the image): the layout a TMA load leaves.
template <int C, int PLANE = PS_PLANE>
__device__ __noinline__ void ps_rows_load_slow(const PsParams& Q, const PsBlk& K, float* s_img, float* s_val, int r0,
                                               int r1, int lane) {
  const int H = Q.p.H, W = Q.p.W;
  const size_t plane = (size_t)H * W;
  const float* img = Q.p.images + (size_t)K.b * 3 * plane;
  const float* val = Q.p.values + (size_t)K.b * C * plane;
  for (int i = r0 * PS_PITCH + lane; i < r1 * PS_PITCH; i += 32) {
    const int t = i / PS_PITCH, cs = i - t * PS_PITCH;
    const int y = K.ys - 2 + t, x = K.x0 - 4 + cs;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    const size_t o = in ? (size_t)y * W + x : 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) s_img[c * PLANE + i] = in ? __ldg(img + c * plane + o) : 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s_val[c * PLANE + i] = in ? __ldg(val + c * plane + o) : 0.f;
  }
}

// Rows [r0, r1) of the tile, in place: image * sqrt(-kc) (sentinel outside the image), values -> probabilities.
// The pass is a latency chain (LDS -> FMUL / MUFU.EX2 / MUFU.RCP -> STS) run by a warp that has nothing else to do, so
// PS_CONV_U groups of four pixels are in flight per lane (loads first, then arithmetic, then stores), the loop carries no
// per-group border logic, and the positions outside the image -- whole rows and whole column ranges of a border tile --
// get their sentinel colour in a short second pass.
#ifndef WSDL_PS_CONV_U
#define WSDL_PS_CONV_U 1
#endif
constexpr int PS_CONV_U = WSDL_PS_CONV_U;

template <int C, int CS, bool SOFTMAX, int PLANE = PS_PLANE>
__device__ __forceinline__ void ps_rows_transform(const PsParams& Q, const PsBlk& K, float* s_img, float* s_val, int r0,
                                                  int r1, int lane) {
  const int H = Q.p.H, W = Q.p.W;
  const float sc = Q.img_scale;
  const int i1 = r1 * PS_Q;
  for (int base0 = r0 * PS_Q; base0 < i1; base0 += 32 * PS_CONV_U) {  // warp-uniform trip count (__syncwarp inside)
    const int base = base0 + lane;
    float4 v[PS_CONV_U][3], u[PS_CONV_U][C];
#pragma unroll
    for (int k = 0; k < PS_CONV_U; ++k) {
      const int so = min(base + 32 * k, i1 - 1) * 4;  // PS_PITCH == 4 * PS_Q; a lane past the end re-reads the last group
#pragma unroll
      for (int c = 0; c < 3; ++c) v[k][c] = *reinterpret_cast<const float4*>(s_img + c * PLANE + so);
#pragma unroll
      for (int c = 0; c < C; ++c) u[k][c] = *reinterpret_cast<const float4*>(s_val + c * PLANE + so);
    }
    __syncwarp();  // every lane holds its groups before anyone overwrites one (the clamped re-reads)
#pragma unroll
    for (int k = 0; k < PS_CONV_U; ++k) {
      const float2 sc2 = make_float2(sc, sc);
#pragma unroll
      for (int c = 0; c < 3; ++c) {  // two-lane FP32: half the issue slots
        const float2 lo = __fmul2_rn(make_float2(v[k][c].x, v[k][c].y), sc2), hi = __fmul2_rn(make_float2(v[k][c].z, v[k][c].w), sc2);
        v[k][c] = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      float w[C][4];
#pragma unroll
      for (int c = 0; c < C; ++c) w[c][0] = u[k][c].x, w[c][1] = u[k][c].y, w[c][2] = u[k][c].z, w[c][3] = u[k][c].w;
      if (CS != C) {  // p0 = 1 / (1 + e^(v1 - v0))
        const float2 l2 = make_float2(LOG2E, LOG2E), one = make_float2(1.f, 1.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 z = __fmul2_rn(__fadd2_rn(make_float2(w[1][2 * h], w[1][2 * h + 1]), make_float2(-w[0][2 * h], -w[0][2 * h + 1])), l2);
          const float2 d = __fadd2_rn(make_float2(ex2_approx(z.x), ex2_approx(z.y)), one);
          w[0][2 * h] = rcp_approx(d.x), w[0][2 * h + 1] = rcp_approx(d.y);
        }
      } else if (SOFTMAX) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float m = w[0][e];
#pragma unroll
          for (int c = 1; c < C; ++c) m = fmaxf(m, w[c][e]);
          float sum = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            w[c][e] = ex2_approx((w[c][e] - m) * LOG2E);
            sum += w[c][e];
          }
          const float inv = rcp_approx(sum);
#pragma unroll
          for (int c = 0; c < C; ++c) w[c][e] *= inv;
        }
      }
      if (base + 32 * k < i1) {
        const int so = (base + 32 * k) * 4;
#pragma unroll
        for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(s_img + c * PLANE + so) = v[k][c];
        if (CS != C || SOFTMAX) {
#pragma unroll
          for (int c = 0; c < CS; ++c)
            *reinterpret_cast<float4*>(s_val + c * PLANE + so) = make_float4(w[c][0], w[c][1], w[c][2], w[c][3]);
        }
      }
    }
  }
  // sentinel colour (channel 0) outside the image: block-uniform tests, a handful of stores on border tiles only
  const int cl = max(0, 4 - K.x0);               // staged columns [0, cl) lie left of the image
  const int cr = min(PS_PITCH, W - (K.x0 - 4));  // staged columns [cr, 68) lie right of it
  const int t_top = min(r1, 2 - K.ys), t_bot = max(r0, H - K.ys + 2);  // rows [r0, t_top) and [t_bot, r1) lie outside
  if (cl > 0 || cr < PS_PITCH || t_top > r0 || t_bot < r1) {
    __syncwarp();
    for (int t = r0 + lane; t < r1; t += 32) {  // a lane per row (a warp converts at most 14)
      for (int x = 0; x < cl; ++x) s_img[t * PS_PITCH + x] = PS_SENTINEL;
      for (int x = cr; x < PS_PITCH; ++x) s_img[t * PS_PITCH + x] = PS_SENTINEL;
    }
    const float4 s4 = make_float4(PS_SENTINEL, PS_SENTINEL, PS_SENTINEL, PS_SENTINEL);
    if (lane < PS_Q) {
      for (int tt = r0; tt < t_top; ++tt) *reinterpret_cast<float4*>(s_img + tt * PS_PITCH + 4 * lane) = s4;
      for (int tt = t_bot; tt < r1; ++tt) *reinterpret_cast<float4*>(s_img + tt * PS_PITCH + 4 * lane) = s4;
    }
  }
}

template <int CS>
__device__ __forceinline__ void ps_zero(float (&X)[8][CS]) {
#pragma unroll
  for (int w = 0; w < 8; ++w)
#pragma unroll
    for (int c = 0; c < CS; ++c) X[w][c] = 0.f;
}

template <int C, bool SOFTMAX>
struct PsCfg {
  static constexpr int CS = (C == 2 && SOFTMAX) ? 1 : C;  // accumulated channels
  static constexpr int CTAS = WSDL_PS_CTAS;               // resident CTAs per SM (<= 128 registers per thread)
  static constexpr size_t smem_floats =
      (size_t)(3 + C) * PS_PLANE + (size_t)(PS_SEGS - 1) * 2 * CS * 64 + 6 * (size_t)CS * PS_CAP + 6 * 2 * 4 + 11 * 5 * 4 + 4;
};

// =====================================================================================================
// Dual kernel: the cut loss on the logits AND the boundary loss on softmax(logits) of the same two-class batch in one
// pass (the composition of BASELINE configs[1] / config 4: LocalNormalizedCutLoss(x, I) and
// ConstrainToBoundaryLossSingle(softmax(x)[b], I[b]) for every image b).  Both losses see the same probability map
// (the cut loss applies its softmax inside, AlternatingDirectionCutLoss.py:78) and differ in the affinity only, so a
// pair shares its colour differences, its squared distance and p(a) - p(b): two exponents from one FFMA chain
// (e_b = ratio e_c + k'), two MUFU, two scatters -- 12 FP32 + 2 MUFU per pair instead of 10 + 13 + 2 in two launches,
// and one tile load, one conversion, one emit.  Two accumulated channels (G of the cut loss, G of the boundary loss)
// over ONE stored probability channel (p1 = 1 - p0 for both).  Output: both loss values and
// d(go_cut L_cut + sum_b go_bnd[b] L_bnd[b]) / d logits.
constexpr unsigned long long PS_SLOT_EMPTY = ~0ull;
__device__ __forceinline__ unsigned long long ps_ld_slot(const unsigned long long* p) {  // from L2, never hoisted
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

struct PsDual {
  const float* grad_out_bnd;  // nullable, B upstream gradients of the per-image boundary losses (cut: Q.p.grad_out)
  float* loss_bnd;            // B floats (cut: Q.p.loss_out, 1 float)
  unsigned long long* slots;  // one per CTA: (cut partial, boundary partial) as two floats, PS_SLOT_EMPTY when unused
  double kappa_bnd;           // 1 / (K H W)
  float kappa_bnd4;           // (float)(4 kappa_bnd)
  float ratio;                // (sigma_cut / sigma_bnd)^2: squared distances are staged for the cut loss
  float ksu_b;                // spatial exponent unit of the boundary loss
  float g1b, g4b, l1g_b;      // gamma, gamma^4, log2(1 + gamma^4) of the boundary loss (cut: 1, 1, 1)
  // Weight tables of the band-column pass, the same for every CTA of a launch: built once on the host (ps_band_tables),
  // copied to shared memory by the border tiles.  fx [6 slots][2][8]: per band slot its two partner columns whose weights
  // differ from the interior's (offset, then 1/2 of Wx(a->b), Wx(b->a), 1 for the cut loss and the same three for the
  // boundary loss); wy [11][5][8]: per row-band row (0..4, H-5..H-1, "any other") and partner row its forward / backward /
  // applied-by-the-march row weights for the cut loss and for the boundary loss.
  float fx[6 * 2 * 8];
  float wy[11 * 5 * 8];
};

// (H >= 10: the row bands of the two image borders do not overlap)
inline void ps_band_tables1(PsParams& Q, int H, int W) {
  const float g1 = Q.g1, g4 = Q.g4;
  for (int slot = 0; slot < 6; ++slot) {
    float w[10];
    const int x = slot < 3 ? slot : W - 6 + slot;
    for (int r = 0; r < 10; ++r) {
      const int xb = x + r % 5 - 2;
      w[r] = (xb < 0 || xb >= W) ? 0.f : (r < 5 ? ps_w1d(x, xb, W, g1, g4) : ps_w1d(xb, x, W, g1, g4));
    }
    for (int u = 0; u < 2; ++u) {
      int found = -1, cnt = 0;
      for (int j = 0; j < 5; ++j) {
        const float g = ps_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
        if (w[j] != 0.f && (w[j] != g || w[5 + j] != g)) {
          if (cnt == u) found = j;
          ++cnt;
        }
      }
      const int j = found < 0 ? 2 : found;  // no such column: the pixel's own, weight 0
      const float live = found < 0 ? 0.f : 0.5f;
      float* fx = Q.fx1 + (slot * 2 + u) * 4;
      memcpy(&fx[0], &j, sizeof(int));
      fx[1] = live * w[j], fx[2] = live * w[5 + j], fx[3] = live * ps_gpow(j < 2 ? 2 - j : j - 2, g1, g4);
    }
  }
  for (int rs = 0; rs < 11; ++rs)
    for (int i = 0; i < 5; ++i) {
      const float g = ps_gpow(i < 2 ? 2 - i : i - 2, g1, g4);
      float* w = Q.wy1 + (rs * 5 + i) * 4;
      w[0] = g, w[1] = g, w[2] = 2.f * g, w[3] = 0.f;
      if (rs < 10) {
        const int zy = rs < 5 ? rs : H - 10 + rs, yb = zy + i - 2;
        const bool in = yb >= 0 && yb < H;
        w[0] = in ? ps_w1d(zy, yb, H, g1, g4) : 0.f;
        w[1] = in ? ps_w1d(yb, zy, H, g1, g4) : 0.f;
        w[2] = in ? 2.f * g * ps_row_mult(zy, yb, H, g4) : 0.f;
      }
    }
}

inline void ps_band_tables(PsDual& D, int H, int W) {
  for (int slot = 0; slot < 6; ++slot) {
    float wc[10], wb[10];
    const int x = slot < 3 ? slot : W - 6 + slot;
    for (int r = 0; r < 10; ++r) {
      const int xb = x + r % 5 - 2;
      const bool in = xb >= 0 && xb < W;
      wc[r] = !in ? 0.f : (r < 5 ? ps_w1d(x, xb, W, 1.f, 1.f) : ps_w1d(xb, x, W, 1.f, 1.f));
      wb[r] = !in ? 0.f : (r < 5 ? ps_w1d(x, xb, W, D.g1b, D.g4b) : ps_w1d(xb, x, W, D.g1b, D.g4b));
    }
    for (int u = 0; u < 2; ++u) {
      int found = -1, cnt = 0;
      for (int j = 0; j < 5; ++j)
        if (wc[j] != 0.f && (wc[j] != 1.f || wc[5 + j] != 1.f)) {
          if (cnt == u) found = j;
          ++cnt;
        }
      const int j = found < 0 ? 2 : found;  // no such column: the pixel's own, weight 0
      const float live = found < 0 ? 0.f : 0.5f;
      float* fx = D.fx + (slot * 2 + u) * 8;
      memcpy(&fx[0], &j, sizeof(int));
      fx[1] = live * wc[j], fx[2] = live * wc[5 + j], fx[3] = live;
      fx[4] = live * wb[j], fx[5] = live * wb[5 + j], fx[6] = live * ps_gpow(j < 2 ? 2 - j : j - 2, D.g1b, D.g4b), fx[7] = 0.f;
    }
  }
  for (int rs = 0; rs < 11; ++rs)
    for (int i = 0; i < 5; ++i) {
      const float g = ps_gpow(i < 2 ? 2 - i : i - 2, D.g1b, D.g4b);
      float* w = D.wy + (rs * 5 + i) * 8;
      w[0] = 1.f, w[1] = 1.f, w[2] = 2.f, w[3] = g, w[4] = g, w[5] = 2.f * g, w[6] = 0.f, w[7] = 0.f;
      if (rs < 10) {
        const int zy = rs < 5 ? rs : H - 10 + rs, yb = zy + i - 2;
        const bool in = yb >= 0 && yb < H;
        w[0] = in ? ps_w1d(zy, yb, H, 1.f, 1.f) : 0.f;
        w[1] = in ? ps_w1d(yb, zy, H, 1.f, 1.f) : 0.f;
        w[2] = in ? 2.f * ps_row_mult(zy, yb, H, 1.f) : 0.f;
        w[3] = in ? ps_w1d(zy, yb, H, D.g1b, D.g4b) : 0.f;
        w[4] = in ? ps_w1d(yb, zy, H, D.g1b, D.g4b) : 0.f;
        w[5] = in ? 2.f * g * ps_row_mult(zy, yb, H, D.g4b) : 0.f;
      }
    }
}

struct PsKsDual {  // per pair class: cut exponent offset, and boundary offset minus ratio * cut offset
  float ca, cb, cc;  // cut: same row, one row down, two rows down (no spatial term: one value per row distance)
  float a1, a4, b0, b1, b4, c0, c1, c4;
};

// ---- packed (f32x2) form of the dual step -------------------------------------------------------------------
// sm_100a has two-lane FP32 instructions (FADD2 / FFMA2, with operand negation and scalar broadcast operands): the
// same lane rate as the scalar ones but HALF the issue slots, and this kernel is bound by issue slots.  Two pairs
// with the same offset whose centres are column neighbours are evaluated per packed instruction: 12 FADD2/FFMA2 +
// 4 MUFU for two pairs instead of 24 + 4.  A packed operand is an (even, odd) register pair, so:
//   * accumulator rows and staged windows are float2 "even pairs" E_k = columns (2k, 2k+1) of the 8-column window,
//     straight from the 128-bit shared-memory loads;
//   * even dx: centres E1, E2 against partners E_(1+dx/2), E_(2+dx/2);
//   * odd dx between rows: the pairs are grouped by their PARTNER column (2..5 = E1, E2 of row t+1 / t+2), their
//     centres are then the "odd pairs" O_k = columns (2k+1, 2k+2) of row t: O1, O2 for dx = -1 and O0, O1 for
//     dx = +1.  Every partner column belongs to exactly one strip, so each pair is still evaluated once;
//   * same row, dx = 1: centres E1, E2 against partners O1, O2.
// Only row t is ever needed at the odd alignment: its 12 odd input pairs are built once per step, and its odd
// accumulators XO (columns 1..6) are folded into the even ones before the row is exchanged and emitted.
// two pairs: centres a (4 planes), partners b; ga/gb = (cut, boundary) accumulators of the centres / partners
__device__ __forceinline__ void ps_pair2(float2 (&ga)[2], float2 (&gb)[2], const float2 (&a)[4], const float2 (&b)[4],
                                         float kc_off, float kb_off, float ratio) {
  const float2 d0 = __fadd2_rn(a[0], f2neg(b[0])), d1 = __fadd2_rn(a[1], f2neg(b[1])), d2 = __fadd2_rn(a[2], f2neg(b[2]));
  const float2 ec =
      __ffma2_rn(f2neg(d2), d2, __ffma2_rn(f2neg(d1), d1, __ffma2_rn(f2neg(d0), d0, make_float2(kc_off, kc_off))));
  const float2 eb = __ffma2_rn(ec, make_float2(ratio, ratio), make_float2(kb_off, kb_off));
  const float2 kc = make_float2(ex2_approx(ec.x), ex2_approx(ec.y));
  const float2 kb = make_float2(ex2_approx(eb.x), ex2_approx(eb.y));
  const float2 dp = __fadd2_rn(a[3], f2neg(b[3]));
  ga[0] = __ffma2_rn(kc, dp, ga[0]);
  gb[0] = __ffma2_rn(f2neg(kc), dp, gb[0]);
  ga[1] = __ffma2_rn(kb, dp, ga[1]);
  gb[1] = __ffma2_rn(f2neg(kb), dp, gb[1]);
}

#define PS_PL(w, k) {w.e[0][k], w.e[1][k], w.e[2][k], w.e[3][k]}

// All 12 forward pairs of the columns of row t.  X, Y, Z: even-pair accumulators [pair][cut, boundary] of rows t, t+1,
// t+2.  On return X also holds the odd-aligned contributions of this step (columns 1..6).
template <int PLANE = PS_PLANE>
__device__ __forceinline__ void ps_step_dual2(float2 (&X)[4][2], float2 (&Y)[4][2], float2 (&Z)[4][2], float (&pc)[4],
                                              const float* s_img, const float* s_p, int off, const PsKsDual& ks, float ratio) {
  PsWinP<1> c;
  ps_loadp<1, PLANE>(c, s_img, s_p, off);
  pc[0] = c.e[3][1].x, pc[1] = c.e[3][1].y, pc[2] = c.e[3][2].x, pc[3] = c.e[3][2].y;
  float2 co[3][4];  // odd pairs of row t: [k][plane] = columns (2k+1, 2k+2)
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int pl = 0; pl < 4; ++pl) co[k][pl] = make_float2(c.e[pl][k].y, c.e[pl][k + 1].x);
  float2 XO[3][2];
#pragma unroll
  for (int k = 0; k < 3; ++k) XO[k][0] = XO[k][1] = make_float2(0.f, 0.f);
  const float2 ce1[4] = PS_PL(c, 1), ce2[4] = PS_PL(c, 2), ce3[4] = PS_PL(c, 3);
  // same row
  ps_pair2(X[1], XO[1], ce1, co[1], ks.ca, ks.a1, ratio);
  ps_pair2(X[2], XO[2], ce2, co[2], ks.ca, ks.a1, ratio);
  ps_pair2(X[1], X[2], ce1, ce2, ks.ca, ks.a4, ratio);
  ps_pair2(X[2], X[3], ce2, ce3, ks.ca, ks.a4, ratio);
#pragma unroll
  for (int r = 1; r <= 2; ++r) {
    float2 (&Yr)[4][2] = r == 1 ? Y : Z;
    const float kc = r == 1 ? ks.cb : ks.cc;
    const float k0 = r == 1 ? ks.b0 : ks.c0, k1 = r == 1 ? ks.b1 : ks.c1, k4 = r == 1 ? ks.b4 : ks.c4;
    PsWinP<1> n;
    ps_loadp<1, PLANE>(n, s_img, s_p, off + r * PS_PITCH);
    const float2 n0[4] = PS_PL(n, 0), n1[4] = PS_PL(n, 1), n2[4] = PS_PL(n, 2), n3[4] = PS_PL(n, 3);
    ps_pair2(X[1], Yr[0], ce1, n0, kc, k4, ratio);     // dx = -2
    ps_pair2(X[2], Yr[1], ce2, n1, kc, k4, ratio);
    ps_pair2(XO[1], Yr[1], co[1], n1, kc, k1, ratio);  // dx = -1: centres 3,4 and 5,6
    ps_pair2(XO[2], Yr[2], co[2], n2, kc, k1, ratio);
    ps_pair2(X[1], Yr[1], ce1, n1, kc, k0, ratio);     // dx = 0
    ps_pair2(X[2], Yr[2], ce2, n2, kc, k0, ratio);
    ps_pair2(XO[0], Yr[1], co[0], n1, kc, k1, ratio);  // dx = +1: centres 1,2 and 3,4
    ps_pair2(XO[1], Yr[2], co[1], n2, kc, k1, ratio);
    ps_pair2(X[1], Yr[2], ce1, n2, kc, k4, ratio);     // dx = +2
    ps_pair2(X[2], Yr[3], ce2, n3, kc, k4, ratio);
  }
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    X[0][ch].y += XO[0][ch].x;
    X[1][ch].x += XO[0][ch].y;
    X[1][ch].y += XO[1][ch].x;
    X[2][ch].x += XO[1][ch].y;
    X[2][ch].y += XO[2][ch].x;
    X[3][ch].x += XO[2][ch].y;
  }
}

}  // namespace wsdl
