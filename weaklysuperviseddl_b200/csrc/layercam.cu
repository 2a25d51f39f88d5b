// LayerCAM -> per-image min-max -> bilinear upsample -> layer fusion -> threshold, for sm_100a.
//
// Replaces the post-backbone half of LayerCAMGenerator.generate
// (reference TraditionalModel/LayerCAM.py:52-76; variant AlternatingDirectionCutLoss.py:262-286)
// and the threshold idiom of PsuedoMasks.py:59-62.
//
// Two kernels per call, both HBM-bound, no tensor cores (nothing here is a contraction):
//
//  A  layercam_channel_sum   reads act/grad (B,C,h,w) once with 128-bit streaming loads and
//     reduces relu(grad*act) over channels.  One CTA owns a (image, layer, pixel tile, channel
//     split).  Channel splits meet through an L2-resident partial buffer and a ticket; the last
//     CTA of a tile applies the outer relu, writes the low-resolution map and folds the tile's
//     min / max into per-(image,layer) atomics (values are >= 0, so uint ordering == float
//     ordering and the result is order independent).  The last tile of an (image, layer)
//     normalises the low-resolution map in place: c -= min; c /= (max(c) + 1e-8) with true
//     IEEE division, exactly the reference's arithmetic (LayerCAM.py:62-67).
//
//  B  layercam_upsample_fuse  each thread owns four adjacent output columns and walks down rows; the
//     horizontally interpolated values of the two low-resolution rows are kept in registers
//     and reused for every output row of the same cell, in torch's rounding order
//     (ATen/native/UpSample.h: horizontal first, then vertical).  Layers are averaged,
//     clamp(0) ** alpha applied and the pixel is thresholded; only the u8 mask (and, on
//     request, the f32 CAM) ever reaches HBM.
#include <cooperative_groups.h>
#include <string.h>

#include "common.cuh"

namespace wsdl {

namespace cg = cooperative_groups;

constexpr int LC_THREADS = 256;
constexpr int LC_UNROLL = 4;

struct LcLayer {
  const void* act;
  const void* grad;
  int C, hw;
  int npg;           // pixel groups (hw / vec)
  int tile_pg;       // pixel groups per tile (<= LC_THREADS)
  int n_tiles;       // tiles per image
  int slices;        // channel slices inside one CTA (LC_THREADS / tile_pg)
  int n_splits;      // channel splits across CTAs
  int ch_per_split;  // channels per split
  int item_begin;    // first blockIdx of this layer
  int ticket_tile_off, ticket_img_off, stats_off;  // u32 offsets into ctrl
  long long low_off;      // float offset into low
  long long partial_off;  // float offset into partial
};

struct LcParams {
  LcLayer L[WSDL_MAX_LAYERS];
  int n_layers, B;
  float alpha;
  int alpha_mode;
  unsigned* ctrl;
  float* low;
  float* partial;
};

template <int DTYPE, int VEC>
__device__ __forceinline__ void load_vec(const char* p, float (&out)[VEC]) {
  if constexpr (DTYPE == WSDL_F32) {
    if constexpr (VEC == 4) {
      uint4 r = ld_stream_u4(p);
      out[0] = __uint_as_float(r.x), out[1] = __uint_as_float(r.y);
      out[2] = __uint_as_float(r.z), out[3] = __uint_as_float(r.w);
    } else {
      static_assert(VEC == 1 || VEC == 4, "f32: vec 1 or 4");
      out[0] = ld_stream_f32(reinterpret_cast<const float*>(p));
    }
  } else {
    if constexpr (VEC == 8) {
      uint4 r = ld_stream_u4(p);
      unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        out[2 * i] = bits16_to_f32<DTYPE>((unsigned short)(w[i] & 0xffffu));
        out[2 * i + 1] = bits16_to_f32<DTYPE>((unsigned short)(w[i] >> 16));
      }
    } else {
      static_assert(VEC == 1 || VEC == 8, "16-bit: vec 1 or 8");
      out[0] = bits16_to_f32<DTYPE>(ld_stream_u16(p));
    }
  }
}

// torch's pow(tensor, python_float) fast paths (ATen/native/cpu/PowKernel.cpp) so that the common
// exponents are bit-identical to the reference; everything else goes through powf.
__device__ __forceinline__ float pow_like_torch(float x, float a) {
  if (a == 1.0f) return x;
  if (a == 2.0f) return x * x;
  if (a == 3.0f) return x * x * x;
  if (a == 0.5f) return sqrtf(x);
  if (a == 0.0f) return 1.0f;
  if (a == -1.0f) return __fdiv_rn(1.0f, x);
  if (a == -0.5f) return __fdiv_rn(1.0f, sqrtf(x));
  if (a == -2.0f) return __fdiv_rn(1.0f, x * x);
  return powf(x, a);
}

struct MinMax {
  float mn, mx;
};

// Block-wide min/max over LC_THREADS threads; result valid in every thread.
__device__ __forceinline__ MinMax block_minmax(float mn, float mx, float* s_red /* >= 16 floats */) {
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // s_red may still be read from a previous call
  if (lane == 0) {
    s_red[warp] = mn;
    s_red[8 + warp] = mx;
  }
  __syncthreads();
  float a = s_red[lane & 7], b = s_red[8 + (lane & 7)];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  }
  return {a, b};
}

template <int DTYPE, int VEC>
__global__ void __launch_bounds__(LC_THREADS) layercam_channel_sum(const __grid_constant__ LcParams P) {
  constexpr int ESIZE = (DTYPE == WSDL_F32) ? 4 : 2;
  __shared__ float s_acc[LC_THREADS * VEC];
  __shared__ float s_red[16];
  __shared__ int s_flag;

  int item = blockIdx.x, l = 0;
  while (l + 1 < P.n_layers && item >= P.L[l + 1].item_begin) ++l;
  const LcLayer& D = P.L[l];
  int r = item - D.item_begin;
  const int split = r % D.n_splits;
  r /= D.n_splits;
  const int tile = r % D.n_tiles;
  const int b = r / D.n_tiles;

  const int tid = threadIdx.x;
  const int pg0 = tile * D.tile_pg;
  const int npg_tile = min(D.tile_pg, D.npg - pg0);
  const int slice = tid / D.tile_pg;
  const int p = tid - slice * D.tile_pg;
  const bool active = (slice < D.slices) && (p < npg_tile);
  const int c_begin = split * D.ch_per_split;
  const int c_end = min(D.C, c_begin + D.ch_per_split);

  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;

  if (active) {
    const size_t plane_bytes = (size_t)D.hw * ESIZE;
    const size_t base = ((size_t)b * D.C * D.hw + (size_t)(pg0 + p) * VEC) * ESIZE;
    const char* pa = reinterpret_cast<const char*>(D.act) + base;
    const char* pg = reinterpret_cast<const char*>(D.grad) + base;
    const int S = D.slices;
    for (int c = c_begin + slice; c < c_end; c += S * LC_UNROLL) {
      float a[LC_UNROLL][VEC], g[LC_UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < LC_UNROLL; ++u) {
        const int cc = c + u * S;
        if (cc < c_end) {
          load_vec<DTYPE, VEC>(pa + (size_t)cc * plane_bytes, a[u]);
          load_vec<DTYPE, VEC>(pg + (size_t)cc * plane_bytes, g[u]);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) a[u][v] = 0.f, g[u][v] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < LC_UNROLL; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += fmaxf(g[u][v] * a[u][v], 0.f);  // relu(grad*act), LayerCAM.py:57
    }
  }

  // ---- channel slices of this CTA -> slice 0 (fixed order: deterministic) ----
  if (D.slices > 1) {
    if (active && slice > 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) s_acc[(slice * D.tile_pg + p) * VEC + v] = acc[v];
    }
    __syncthreads();
    if (active && slice == 0) {
      for (int s = 1; s < D.slices; ++s)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += s_acc[(s * D.tile_pg + p) * VEC + v];
    }
  }
  const bool owner = active && slice == 0;  // holds the tile's channel sum over [c_begin, c_end)

  // ---- channel splits across CTAs -> last arriver ----
  if (D.n_splits > 1) {
    const size_t tile_px = (size_t)D.tile_pg * VEC;
    float* part = P.partial + D.partial_off + ((size_t)(b * D.n_tiles + tile) * D.n_splits) * tile_px;
    if (owner) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) __stcg(part + (size_t)split * tile_px + p * VEC + v, acc[v]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned t = atomicAdd(P.ctrl + D.ticket_tile_off + b * D.n_tiles + tile, 1u);
      s_flag = (t == (unsigned)D.n_splits - 1u);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (owner) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
      for (int s = 0; s < D.n_splits; ++s)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += ld_cg_f32(part + (size_t)s * tile_px + p * VEC + v);
    }
  }

  // ---- tile finalise: outer relu (LayerCAM.py:59), low-res write, min/max ----
  float* low = P.low + D.low_off + (size_t)b * D.hw;
  float tmn = __int_as_float(0x7f800000), tmx = 0.f;
  if (owner) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float x = fmaxf(acc[v], 0.f);
      __stcg(low + (size_t)(pg0 + p) * VEC + v, x);
      tmn = fminf(tmn, x);
      tmx = fmaxf(tmx, x);
    }
  }
  MinMax mm = block_minmax(tmn, tmx, s_red);
  unsigned* stats = P.ctrl + D.stats_off + 2 * b;
  bool last_tile = true;
  if (D.n_tiles > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      atomicMax(stats + 0, ~(__float_as_uint(mm.mn) & 0x7fffffffu));  // min via max of complement
      atomicMax(stats + 1, __float_as_uint(mm.mx) & 0x7fffffffu);
      __threadfence();
      const unsigned t = atomicAdd(P.ctrl + D.ticket_img_off + b, 1u);
      s_flag = (t == (unsigned)D.n_tiles - 1u);
    }
    __syncthreads();
    last_tile = s_flag;
    if (!last_tile) return;
    __threadfence();
    mm.mn = __uint_as_float(~ld_cg_u32(stats + 0));
    mm.mx = __uint_as_float(ld_cg_u32(stats + 1));
  }

  // ---- image finalise: per-image min-max normalisation, in place (LayerCAM.py:62-67) ----
  const float mn = mm.mn;
  const float den = (mm.mx - mn) + 1e-8f;  // max is taken after the subtraction in the reference
  const bool single = (D.n_tiles == 1);    // then this thread's values are still in registers
  if (P.alpha_mode == 0) {
    if (single) {
      if (owner) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          __stcg(low + (size_t)(pg0 + p) * VEC + v, __fdiv_rn(fmaxf(acc[v], 0.f) - mn, den));
      }
    } else {
      for (int i = tid; i < D.hw; i += LC_THREADS) __stcg(low + i, __fdiv_rn(ld_cg_f32(low + i) - mn, den));
    }
  } else {
    // variant (AlternatingDirectionCutLoss.py:271-279): normalise, ** alpha, normalise again
    float mn2 = __int_as_float(0x7f800000), mx2 = -__int_as_float(0x7f800000);
    for (int i = tid; i < D.hw; i += LC_THREADS) {
      const float x = pow_like_torch(__fdiv_rn(ld_cg_f32(low + i) - mn, den), P.alpha);
      __stcg(low + i, x);
      mn2 = fminf(mn2, x);
      mx2 = fmaxf(mx2, x);
    }
    MinMax m2 = block_minmax(mn2, mx2, s_red);
    const float den2 = (m2.mx - m2.mn) + 1e-8f;
    for (int i = tid; i < D.hw; i += LC_THREADS) __stcg(low + i, __fdiv_rn(ld_cg_f32(low + i) - m2.mn, den2));
  }
}

// ------------------------------------------------------------------------------------------
// Kernel B
// ------------------------------------------------------------------------------------------
constexpr int UP_THREADS = 128;
constexpr int UP_ROWS = 32;

struct UpLayer {
  const float* low;  // (B, h, w) normalised
  int h, w;
  float scale_y, scale_x;  // float(in) / float(out), as area_pixel_compute_scale
};

struct UpParams {
  UpLayer L[WSDL_MAX_LAYERS];
  int n_layers, B, out_h, out_w;
  float alpha;
  int alpha_mode;
  float thresh, band;
  float inv_layers;  // 1/L when L is a power of two (exact), else 0 -> true division
  float n_layers_f;
  float* cam_out;
  uint8_t* mask_out;
  unsigned long long* near_count;
  int vec_ok;  // out_w % 4 == 0, cam_out 16-byte and mask_out 4-byte aligned: one 128-bit / 32-bit store per row
};

struct Tap {
  int i0, i1;
  float l0, l1;
};

// torch align_corners=False source index + guard (ATen/native/UpSample.h:259-300,469-475).
__device__ __forceinline__ Tap make_tap(float scale, int dst, int in_size) {
  float src = __fsub_rn(__fmul_rn(scale, (float)dst + 0.5f), 0.5f);
  src = src < 0.f ? 0.f : src;
  int i0 = min((int)floorf(src), in_size - 1);
  Tap t;
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  t.l0 = 1.f - t.l1;
  return t;
}

__device__ __forceinline__ float lerp2(float l0, float a, float l1, float b) {
  return __fadd_rn(__fmul_rn(l0, a), __fmul_rn(l1, b));  // no FMA contraction: matches the eager op order
}

constexpr int UP_COLS = 4;  // output columns per thread: one 32-bit mask store (128-bit CAM store) per row

template <int NL>  // number of layers, 1..4: taps and row caches live in registers
__global__ void __launch_bounds__(UP_THREADS) layercam_upsample_fuse(const __grid_constant__ UpParams P) {
  __shared__ Tap s_ytap[NL][UP_ROWS];
  const int tid = threadIdx.x;
  const int x0 = (blockIdx.x * UP_THREADS + tid) * UP_COLS;
  const int y_begin = blockIdx.y * UP_ROWS;
  const int b = blockIdx.z;
  const int rows = min(UP_ROWS, P.out_h - y_begin);

  for (int i = tid; i < NL * UP_ROWS; i += UP_THREADS) {
    const int l = i / UP_ROWS, rr = i - l * UP_ROWS;
    if (rr < rows) s_ytap[l][rr] = make_tap(P.L[l].scale_y, y_begin + rr, P.L[l].h);
  }
  __syncthreads();
  // no early return: lanes past the right edge stay for the warp reduction of `near` (they own no column)
  const int ncol = max(0, min(UP_COLS, P.out_w - x0));
  const bool vec = (ncol == UP_COLS) && P.vec_ok;  // aligned 4-column group

  Tap xt[NL][UP_COLS];
  const float* base[NL];
  int cur0[NL], cur1[NL];
  float h0[NL][UP_COLS], h1[NL][UP_COLS];
#pragma unroll
  for (int l = 0; l < NL; ++l) {
#pragma unroll
    for (int c = 0; c < UP_COLS; ++c) {
      xt[l][c] = make_tap(P.L[l].scale_x, min(x0 + c, P.out_w - 1), P.L[l].w);
      h0[l][c] = h1[l][c] = 0.f;
    }
    base[l] = P.L[l].low + (size_t)b * P.L[l].h * P.L[l].w;
    cur0[l] = cur1[l] = -1;
  }

  unsigned near = 0;
  const size_t out_base = ((size_t)b * P.out_h + y_begin) * P.out_w + x0;
  // launch-uniform choices, hoisted out of the pixel loop
  const bool mean_by_mul = P.inv_layers != 0.f;
  const int shape = P.alpha_mode != 0 ? 0 : (P.alpha == 1.0f ? 1 : 2);  // 0: as is, 1: clamp only, 2: clamp + pow
  const bool count_near = P.near_count != nullptr;
  for (int rr = 0; rr < rows; ++rr) {
    float s[UP_COLS];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const Tap yt = s_ytap[l][rr];
      if (yt.i0 != cur0[l] || yt.i1 != cur1[l]) {  // CTA-uniform
        const int w = P.L[l].w;
        if (yt.i0 == cur1[l]) {
#pragma unroll
          for (int c = 0; c < UP_COLS; ++c) h0[l][c] = h1[l][c];
        } else {
          const float* r0 = base[l] + (size_t)yt.i0 * w;
#pragma unroll
          for (int c = 0; c < UP_COLS; ++c)
            h0[l][c] = lerp2(xt[l][c].l0, __ldg(r0 + xt[l][c].i0), xt[l][c].l1, __ldg(r0 + xt[l][c].i1));
        }
        if (yt.i1 == yt.i0) {
#pragma unroll
          for (int c = 0; c < UP_COLS; ++c) h1[l][c] = h0[l][c];
        } else {
          const float* r1 = base[l] + (size_t)yt.i1 * w;
#pragma unroll
          for (int c = 0; c < UP_COLS; ++c)
            h1[l][c] = lerp2(xt[l][c].l0, __ldg(r1 + xt[l][c].i0), xt[l][c].l1, __ldg(r1 + xt[l][c].i1));
        }
        cur0[l] = yt.i0;
        cur1[l] = yt.i1;
      }
#pragma unroll
      for (int c = 0; c < UP_COLS; ++c) {
        const float v = lerp2(yt.l0, h0[l][c], yt.l1, h1[l][c]);
        s[c] = (l == 0) ? v : s[c] + v;  // python sum(): 0 + cam_0 + cam_1 ... (LayerCAM.py:74)
      }
    }
    float cam[UP_COLS];
    unsigned bits = 0;
#pragma unroll
    for (int c = 0; c < UP_COLS; ++c) {
      float v = mean_by_mul ? s[c] * P.inv_layers : __fdiv_rn(s[c], P.n_layers_f);
      if (shape == 1) v = fmaxf(v, 0.f);                              // clamp(0) ** 1
      else if (shape == 2) v = pow_like_torch(fmaxf(v, 0.f), P.alpha);  // LayerCAM.py:76
      cam[c] = v;
      if (v >= P.thresh && v > 0.f) bits |= 1u << (8 * c);  // PsuedoMasks.py:60-62
      if (count_near && c < ncol) near += (fabsf(v - P.thresh) < P.band) ? 1u : 0u;
    }
    if (ncol == 0) continue;
    const size_t o = out_base + (size_t)rr * P.out_w;
    if (vec) {
      if (P.cam_out) *reinterpret_cast<float4*>(P.cam_out + o) = make_float4(cam[0], cam[1], cam[2], cam[3]);
      if (P.mask_out) *reinterpret_cast<unsigned*>(P.mask_out + o) = bits;
    } else {
#pragma unroll
      for (int c = 0; c < UP_COLS; ++c)
        if (c < ncol) {
          if (P.cam_out) P.cam_out[o + c] = cam[c];
          if (P.mask_out) P.mask_out[o + c] = (uint8_t)((bits >> (8 * c)) & 1u);
        }
    }
  }
  if (P.near_count) {
    near = warp_sum_u32(near);
    if ((tid & 31) == 0 && near) atomicAdd(P.near_count, (unsigned long long)near);
  }
}

// Generic fallback for 5..8 layers: no register caches, four taps per pixel per layer.
__global__ void __launch_bounds__(UP_THREADS) layercam_upsample_fuse_generic(const __grid_constant__ UpParams P) {
  const int tid = threadIdx.x;
  const int x = blockIdx.x * UP_THREADS + tid;
  const int y_begin = blockIdx.y * UP_ROWS;
  const int b = blockIdx.z;
  const int rows = min(UP_ROWS, P.out_h - y_begin);
  const bool in_x = x < P.out_w;
  const int xc = in_x ? x : P.out_w - 1;
  unsigned near = 0;
  for (int rr = 0; rr < rows; ++rr) {
    const int y = y_begin + rr;
    float s = 0.f;
    for (int l = 0; l < P.n_layers; ++l) {
      const Tap yt = make_tap(P.L[l].scale_y, y, P.L[l].h);
      const Tap xt = make_tap(P.L[l].scale_x, xc, P.L[l].w);
      const float* base = P.L[l].low + (size_t)b * P.L[l].h * P.L[l].w;
      const float* r0 = base + (size_t)yt.i0 * P.L[l].w;
      const float* r1 = base + (size_t)yt.i1 * P.L[l].w;
      const float a = lerp2(xt.l0, __ldg(r0 + xt.i0), xt.l1, __ldg(r0 + xt.i1));
      const float c = lerp2(xt.l0, __ldg(r1 + xt.i0), xt.l1, __ldg(r1 + xt.i1));
      const float v = lerp2(yt.l0, a, yt.l1, c);
      s = (l == 0) ? v : s + v;
    }
    float cam = (P.inv_layers != 0.f) ? s * P.inv_layers : __fdiv_rn(s, P.n_layers_f);
    if (P.alpha_mode == 0) cam = pow_like_torch(fmaxf(cam, 0.f), P.alpha);
    if (in_x) {
      const size_t o = ((size_t)b * P.out_h + y) * P.out_w + x;
      if (P.cam_out) P.cam_out[o] = cam;
      if (P.mask_out) P.mask_out[o] = (cam >= P.thresh && cam > 0.f) ? 1 : 0;
      near += (fabsf(cam - P.thresh) < P.band) ? 1u : 0u;
    }
  }
  if (P.near_count) {
    near = warp_sum_u32(near);
    if ((tid & 31) == 0 && near) atomicAdd(P.near_count, (unsigned long long)near);
  }
}

__global__ void threshold_mask_kernel(const float* __restrict__ cam, size_t n, float thresh, float band,
                                      uint8_t* __restrict__ mask, unsigned long long* near_count) {
  unsigned near = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = cam[i];
    if (mask) mask[i] = (c >= thresh && c > 0.f) ? 1 : 0;
    near += (fabsf(c - thresh) < band) ? 1u : 0u;
  }
  if (near_count) {
    near = warp_sum_u32(near);
    if ((threadIdx.x & 31) == 0 && near) atomicAdd(near_count, (unsigned long long)near);
  }
}

// ------------------------------------------------------------------------------------------
// Small batches: ONE launch, one thread-block cluster of 16 CTAs per image.
//
// The reference calls generate() one image at a time (LayerCAM.py:38) and BASELINE config 1 is B = 8: a few MB of
// hooks per call, where the two-kernel path above spends its time on launches, the control memset and the tickets
// through L2 (35 us for 6 us of HBM traffic at B = 8).  Here the 16 CTAs of a cluster share an image: they split the
// channels of its layers, meet through DISTRIBUTED SHARED MEMORY (each layer's leader CTA adds its peers' partial maps
// in rank order, applies the outer ReLU and the per-image min-max), every CTA copies the finished low-resolution maps
// from the leaders, and each upsamples / fuses / thresholds a sixteenth of the output rows.  No workspace, no memset,
// no atomics except the near-threshold counter.
// ------------------------------------------------------------------------------------------
constexpr int SC_CLUSTER = 16;
constexpr int SC_MAX_LOW = 4096;  // floats of all low-resolution maps of one image (16 KB)
constexpr int SC_MAX_HW = 1024;   // pixels of one low-resolution map
// channels in flight per thread (template SC_UNROLL): 16 SMs must pull a whole image, so bytes in flight decide (Little's
// law).  8 for up to 8 images (13.4 vs 15.1 us at B = 1, 17.3 vs 18.1 us at B = 8), 4 beyond (23 vs 31 us at B = 16: two
// waves of clusters).

struct ScParams {
  const void* act[4];
  const void* grad[4];
  int C[4], h[4], w[4], hw[4], low_off[4];
  int first[4], count[4], ch_per_cta[4];  // the cluster ranks that work on a layer, channels each of them takes
  int rank_layer[SC_CLUSTER];
  float scale_y[4], scale_x[4];
  int n_layers, B, out_h, out_w, rows_per_cta;
  float alpha;
  int alpha_mode;
  float thresh, band, inv_layers, n_layers_f;
  float* cam_out;
  uint8_t* mask_out;
  unsigned long long* near_count;
  int vec_ok;
};

template <int DTYPE, int VEC, int NL, int SC_UNROLL>
__global__ void __launch_bounds__(LC_THREADS) layercam_cluster_kernel(const __grid_constant__ ScParams P) {
  constexpr int ESIZE = (DTYPE == WSDL_F32) ? 4 : 2;
  __shared__ __align__(16) float s_acc[LC_THREADS * VEC];
  __shared__ __align__(16) float s_part[SC_MAX_HW];  // this CTA's channel-partial map; at a leader then the finished map
  __shared__ __align__(16) float s_low[SC_MAX_LOW];  // all finished maps of the image
  __shared__ float s_red[16];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.x / SC_CLUSTER;
  const int tid = threadIdx.x;
  const int l = P.rank_layer[rank];
  const int hw = P.hw[l], npg = hw / VEC;
  const int slices = min(LC_THREADS / npg, P.ch_per_cta[l]);
  const int slice = tid / npg, p = tid - slice * npg;
  const bool active = slice < slices;

  // ---- channel sum of this CTA's share: relu(grad * act), LayerCAM.py:57 ----
  float acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  if (active) {
    const int j = rank - P.first[l];
    const int c_begin = j * P.ch_per_cta[l], c_end = min(P.C[l], c_begin + P.ch_per_cta[l]);
    const size_t plane_bytes = (size_t)hw * ESIZE;
    const size_t base = ((size_t)b * P.C[l] * hw + (size_t)p * VEC) * ESIZE;
    const char* pa = reinterpret_cast<const char*>(P.act[l]) + base;
    const char* pg = reinterpret_cast<const char*>(P.grad[l]) + base;
    for (int c = c_begin + slice; c < c_end; c += slices * SC_UNROLL) {
      float a[SC_UNROLL][VEC], g[SC_UNROLL][VEC];
#pragma unroll
      for (int u = 0; u < SC_UNROLL; ++u) {
        const int cc = c + u * slices;
        if (cc < c_end) {
          load_vec<DTYPE, VEC>(pa + (size_t)cc * plane_bytes, a[u]);
          load_vec<DTYPE, VEC>(pg + (size_t)cc * plane_bytes, g[u]);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) a[u][v] = 0.f, g[u][v] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < SC_UNROLL; ++u)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] += fmaxf(g[u][v] * a[u][v], 0.f);
    }
  }
  if (active && slice > 0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) s_acc[(slice * npg + p) * VEC + v] = acc[v];
  }
  __syncthreads();
  if (active && slice == 0) {  // slices of this CTA in a fixed order
    for (int s = 1; s < slices; ++s)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] += s_acc[(s * npg + p) * VEC + v];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s_part[p * VEC + v] = acc[v];
  }
  cluster.sync();  // every CTA's partial map is in its shared memory

  // ---- the layer's leader: peers in rank order, outer ReLU (LayerCAM.py:59), per-image min-max (:62-67) ----
  if (rank == P.first[l]) {
    float x[4], mn = __int_as_float(0x7f800000), mx = 0.f;  // up to 4 pixels per thread (hw <= 1024)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = tid + q * LC_THREADS;
      x[q] = 0.f;
      if (i < hw) {
        float sum = 0.f;
        for (int j = 0; j < P.count[l]; ++j) sum += cluster.map_shared_rank(s_part, rank + j)[i];
        x[q] = fmaxf(sum, 0.f);
        mn = fminf(mn, x[q]), mx = fmaxf(mx, x[q]);
      }
    }
    // (no cluster barrier here: the peers do nothing but wait for the finished maps, and a leader thread overwrites only
    // the entries of its own partial map that it has read itself)
    MinMax mm = block_minmax(mn, mx, s_red);
    const float den = (mm.mx - mm.mn) + 1e-8f;  // max is taken after the subtraction in the reference
    if (P.alpha_mode == 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid + q * LC_THREADS;
        if (i < hw) s_part[i] = __fdiv_rn(x[q] - mm.mn, den);
      }
    } else {  // variant (AlternatingDirectionCutLoss.py:271-279): normalise, ** alpha, normalise again
      float mn2 = __int_as_float(0x7f800000), mx2 = -__int_as_float(0x7f800000);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid + q * LC_THREADS;
        if (i < hw) {
          x[q] = pow_like_torch(__fdiv_rn(x[q] - mm.mn, den), P.alpha);
          mn2 = fminf(mn2, x[q]), mx2 = fmaxf(mx2, x[q]);
        }
      }
      MinMax m2 = block_minmax(mn2, mx2, s_red);
      const float den2 = (m2.mx - m2.mn) + 1e-8f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = tid + q * LC_THREADS;
        if (i < hw) s_part[i] = __fdiv_rn(x[q] - m2.mn, den2);
      }
    }
  }
  cluster.sync();  // the leaders hold the finished maps (and have read every peer's partial map)

  // ---- every CTA takes a copy of all maps of the image ----
#pragma unroll
  for (int k = 0; k < NL; ++k) {
    const float* src = cluster.map_shared_rank(s_part, P.first[k]);
    for (int i = tid; i < P.hw[k]; i += LC_THREADS) s_low[P.low_off[k] + i] = src[i];
  }
  cluster.sync();  // nobody's shared memory is needed by a peer any more

  // ---- upsample + fuse + threshold this CTA's rows (LayerCAM.py:69-76, PsuedoMasks.py:59-62) ----
  const int groups_x = (P.out_w + UP_COLS - 1) / UP_COLS;   // threads across a row, four columns each
  const int lanes_y = LC_THREADS / groups_x;                // row groups of this CTA
  const int gx = tid % groups_x, gy = tid / groups_x;
  const int y_begin = rank * P.rows_per_cta, y_end = min(P.out_h, y_begin + P.rows_per_cta);
  unsigned near = 0;
  if (gy < lanes_y) {
    const int x0 = gx * UP_COLS;
    const int ncol = max(0, min(UP_COLS, P.out_w - x0));
    const bool vec = (ncol == UP_COLS) && P.vec_ok;
    Tap xt[NL][UP_COLS];
    int cur0[NL], cur1[NL];
    float h0[NL][UP_COLS], h1[NL][UP_COLS];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
#pragma unroll
      for (int c = 0; c < UP_COLS; ++c) {
        xt[k][c] = make_tap(P.scale_x[k], min(x0 + c, P.out_w - 1), P.w[k]);
        h0[k][c] = h1[k][c] = 0.f;
      }
      cur0[k] = cur1[k] = -1;
    }
    const bool mean_by_mul = P.inv_layers != 0.f;
    const int shape = P.alpha_mode != 0 ? 0 : (P.alpha == 1.0f ? 1 : 2);
    const bool count_near = P.near_count != nullptr;
    // consecutive rows per thread, so that the horizontal interpolations of a low-resolution row pair are reused
    const int per = (y_end - y_begin + lanes_y - 1) / lanes_y;
    const int ya = y_begin + gy * per, yb = min(y_end, ya + per);
    for (int y = ya; y < yb; ++y) {
      float sacc[UP_COLS];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        const Tap yt = make_tap(P.scale_y[k], y, P.h[k]);
        if (yt.i0 != cur0[k] || yt.i1 != cur1[k]) {
          const int w = P.w[k];
          const float* basek = s_low + P.low_off[k];
          if (yt.i0 == cur1[k]) {
#pragma unroll
            for (int c = 0; c < UP_COLS; ++c) h0[k][c] = h1[k][c];
          } else {
            const float* r0 = basek + yt.i0 * w;
#pragma unroll
            for (int c = 0; c < UP_COLS; ++c) h0[k][c] = lerp2(xt[k][c].l0, r0[xt[k][c].i0], xt[k][c].l1, r0[xt[k][c].i1]);
          }
          if (yt.i1 == yt.i0) {
#pragma unroll
            for (int c = 0; c < UP_COLS; ++c) h1[k][c] = h0[k][c];
          } else {
            const float* r1 = basek + yt.i1 * w;
#pragma unroll
            for (int c = 0; c < UP_COLS; ++c) h1[k][c] = lerp2(xt[k][c].l0, r1[xt[k][c].i0], xt[k][c].l1, r1[xt[k][c].i1]);
          }
          cur0[k] = yt.i0, cur1[k] = yt.i1;
        }
#pragma unroll
        for (int c = 0; c < UP_COLS; ++c) {
          const float v = lerp2(yt.l0, h0[k][c], yt.l1, h1[k][c]);
          sacc[c] = (k == 0) ? v : sacc[c] + v;  // python sum(): 0 + cam_0 + cam_1 ... (LayerCAM.py:74)
        }
      }
      float cam[UP_COLS];
      unsigned bits = 0;
#pragma unroll
      for (int c = 0; c < UP_COLS; ++c) {
        float v = mean_by_mul ? sacc[c] * P.inv_layers : __fdiv_rn(sacc[c], P.n_layers_f);
        if (shape == 1) v = fmaxf(v, 0.f);
        else if (shape == 2) v = pow_like_torch(fmaxf(v, 0.f), P.alpha);
        cam[c] = v;
        if (v >= P.thresh && v > 0.f) bits |= 1u << (8 * c);
        if (count_near && c < ncol) near += (fabsf(v - P.thresh) < P.band) ? 1u : 0u;
      }
      if (ncol == 0) continue;
      const size_t o = ((size_t)b * P.out_h + y) * P.out_w + x0;
      if (vec) {
        if (P.cam_out) *reinterpret_cast<float4*>(P.cam_out + o) = make_float4(cam[0], cam[1], cam[2], cam[3]);
        if (P.mask_out) *reinterpret_cast<unsigned*>(P.mask_out + o) = bits;
      } else {
#pragma unroll
        for (int c = 0; c < UP_COLS; ++c)
          if (c < ncol) {
            if (P.cam_out) P.cam_out[o + c] = cam[c];
            if (P.mask_out) P.mask_out[o + c] = (uint8_t)((bits >> (8 * c)) & 1u);
          }
      }
    }
  }
  if (P.near_count) {
    near = warp_sum_u32(near);
    if ((tid & 31) == 0 && near) atomicAdd(P.near_count, (unsigned long long)near);
  }
}

// ------------------------------------------------------------------------------------------
// host side: work decomposition ("plan"), shared by the workspace query and the launcher
// ------------------------------------------------------------------------------------------
struct LcPlan {
  LcParams P;
  int vec;
  int total_items;
  size_t ctrl_u32, low_f32, partial_f32;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int lc_plan(const int* C, const int* h, const int* w, int n_layers, int B, int dtype, LcPlan* out) {
  if (!C || !h || !w) return WSDL_E_NULL;
  if (n_layers < 1 || n_layers > WSDL_MAX_LAYERS || B < 1) return WSDL_E_SHAPE;
  if (dtype != WSDL_F32 && dtype != WSDL_BF16 && dtype != WSDL_F16) return WSDL_E_DTYPE;
  const int esize = dtype == WSDL_F32 ? 4 : 2;
  int vec = 16 / esize;  // 128-bit loads when every layer's plane is a multiple of it
  for (int l = 0; l < n_layers; ++l) {
    if (C[l] < 1 || h[l] < 1 || w[l] < 1) return WSDL_E_SHAPE;
    if ((long long)h[l] * w[l] > (1LL << 30)) return WSDL_E_SHAPE;
    if (((long long)h[l] * w[l]) % vec != 0) vec = 1;
  }
  LcPlan& pl = *out;
  pl.vec = vec;
  size_t ctrl = 0, low = 0, partial = 0;
  long long items = 0;
  // first pass: geometry with a bytes-per-CTA target
  const size_t target_bytes = 512u << 10;  // act+grad bytes one CTA streams
  for (int l = 0; l < n_layers; ++l) {
    LcLayer& D = pl.P.L[l];
    D.C = C[l];
    D.hw = h[l] * w[l];
    D.npg = D.hw / vec;
    D.tile_pg = D.npg < LC_THREADS ? D.npg : LC_THREADS;
    D.n_tiles = (D.npg + D.tile_pg - 1) / D.tile_pg;
    D.slices = LC_THREADS / D.tile_pg;
    if (D.slices > D.C) D.slices = D.C;
    const size_t bytes_per_ch = (size_t)D.tile_pg * vec * esize * 2;
    int want_ch = (int)(target_bytes / bytes_per_ch);
    const int min_ch = D.slices * LC_UNROLL;
    if (want_ch < min_ch) want_ch = min_ch;
    D.n_splits = (D.C + want_ch - 1) / want_ch;
    items += (long long)B * D.n_tiles * D.n_splits;
  }
  // small batches: split channels further until the grid covers the machine a few times over
  const long long min_items = 4LL * WSDL_NUM_SMS;
  if (items < min_items) {
    const int f = (int)((min_items + items - 1) / items);
    for (int l = 0; l < n_layers; ++l) {
      LcLayer& D = pl.P.L[l];
      const int min_ch = D.slices * LC_UNROLL;
      int max_splits = D.C / min_ch;
      if (max_splits < 1) max_splits = 1;
      int ns = D.n_splits * f;
      D.n_splits = ns > max_splits ? max_splits : ns;
    }
  }
  items = 0;
  for (int l = 0; l < n_layers; ++l) {
    LcLayer& D = pl.P.L[l];
    int cps = (D.C + D.n_splits - 1) / D.n_splits;
    const int q = D.slices * LC_UNROLL;  // keep every slice's unrolled loop full
    cps = (cps + q - 1) / q * q;
    D.ch_per_split = cps;
    D.n_splits = (D.C + cps - 1) / cps;
    D.item_begin = (int)items;
    items += (long long)B * D.n_tiles * D.n_splits;
    if (items > 0x7fffffffLL) return WSDL_E_SHAPE;
    D.ticket_tile_off = (int)ctrl;
    ctrl += (size_t)B * D.n_tiles;
    D.ticket_img_off = (int)ctrl;
    ctrl += (size_t)B;
    D.stats_off = (int)ctrl;
    ctrl += (size_t)B * 2;
    D.low_off = (long long)low;
    low += align_up((size_t)B * D.hw, 4);
    D.partial_off = (long long)partial;
    if (D.n_splits > 1) partial += align_up((size_t)B * D.n_tiles * D.n_splits * D.tile_pg * vec, 4);
  }
  pl.total_items = (int)items;
  pl.ctrl_u32 = align_up(ctrl, 64);
  pl.low_f32 = low;
  pl.partial_f32 = partial;
  pl.P.n_layers = n_layers;
  pl.P.B = B;
  return 0;
}

static size_t lc_workspace_bytes(const LcPlan& pl) { return (pl.ctrl_u32 + pl.low_f32 + pl.partial_f32) * 4 + 256; }

template <int DTYPE>
static void launch_channel_sum(const LcPlan& pl, cudaStream_t s) {
  constexpr int WIDE = DTYPE == WSDL_F32 ? 4 : 8;
  if (pl.vec == WIDE)
    layercam_channel_sum<DTYPE, WIDE><<<pl.total_items, LC_THREADS, 0, s>>>(pl.P);
  else
    layercam_channel_sum<DTYPE, 1><<<pl.total_items, LC_THREADS, 0, s>>>(pl.P);
}

template <int DTYPE, int VEC, int UNROLL>
static int sc_launch_t(const ScParams& P, cudaStream_t s, bool query_only) {
  void (*kern)(ScParams) = nullptr;
  switch (P.n_layers) {
    case 1: kern = layercam_cluster_kernel<DTYPE, VEC, 1, UNROLL>; break;
    case 2: kern = layercam_cluster_kernel<DTYPE, VEC, 2, UNROLL>; break;
    case 3: kern = layercam_cluster_kernel<DTYPE, VEC, 3, UNROLL>; break;
    default: kern = layercam_cluster_kernel<DTYPE, VEC, 4, UNROLL>; break;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(SC_CLUSTER * P.B), cfg.blockDim = dim3(LC_THREADS), cfg.dynamicSmemBytes = 0, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SC_CLUSTER, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    (void)cudaGetLastError();
    return 1;
  }
  if (query_only) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) {
      (void)cudaGetLastError();
      return 1;
    }
    return 0;
  }
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
  return e == cudaSuccess ? 0 : (int)e;
}

// 1: not this path's case (the caller takes the two-kernel path), 0: launched, else a CUDA error
static int sc_launch(const void* const* act, const void* const* grad, const int* C, const int* h, const int* w, int n_layers,
                     int B, int dtype, int out_h, int out_w, float alpha, int alpha_mode, float thresh, float band,
                     float* cam_out, uint8_t* mask_out, unsigned long long* near_count, cudaStream_t s) {
  static const int off = WSDL_TUNE_INT("WSDL_LAYERCAM_NO_CLUSTER", 0);
  if (off || n_layers > 4 || B > 16 || out_w > UP_COLS * LC_THREADS || out_w < 1) return 1;
  const int vecw = dtype == WSDL_F32 ? 4 : 8, esize = dtype == WSDL_F32 ? 4 : 2;
  ScParams P;
  memset(&P, 0, sizeof(P));
  long long total_low = 0, work[4], work_sum = 0, bytes = 0;
  for (int l = 0; l < n_layers; ++l) {
    const long long hw = (long long)h[l] * w[l];
    if (hw > SC_MAX_HW || (hw % vecw) != 0 || hw / vecw > LC_THREADS) return 1;
    if (((uintptr_t)act[l] % 16) || ((uintptr_t)grad[l] % 16)) return 1;
    P.act[l] = act[l], P.grad[l] = grad[l], P.C[l] = C[l], P.h[l] = h[l], P.w[l] = w[l], P.hw[l] = (int)hw;
    P.low_off[l] = (int)total_low;
    total_low += hw;
    work[l] = (long long)C[l] * hw;
    work_sum += work[l];
    bytes += 2LL * C[l] * hw * esize;
    P.scale_y[l] = (float)h[l] / (float)out_h;
    P.scale_x[l] = (float)w[l] / (float)out_w;
  }
  if (total_low > SC_MAX_LOW || n_layers > SC_CLUSTER) return 1;
  // one cluster = 16 SMs of one GPC: fine for a few MB per image; bigger images and batches are better off on the
  // streaming path, which spreads every image over the whole machine and reaches the HBM roofline there
  if (bytes > (12LL << 20) || bytes * B > (96LL << 20)) return 1;
  // ranks per layer in proportion to the bytes it streams, at least one each
  int ranks[4], used = 0;
  for (int l = 0; l < n_layers; ++l) {
    ranks[l] = (int)((work[l] * SC_CLUSTER) / work_sum);
    if (ranks[l] < 1) ranks[l] = 1;
    used += ranks[l];
  }
  while (used > SC_CLUSTER) {  // take from the layer with the least work per rank
    int k = -1;
    for (int l = 0; l < n_layers; ++l)
      if (ranks[l] > 1 && (k < 0 || work[l] * ranks[k] < work[k] * ranks[l])) k = l;
    if (k < 0) return 1;
    --ranks[k], --used;
  }
  while (used < SC_CLUSTER) {  // give to the layer with the most work per rank
    int k = 0;
    for (int l = 1; l < n_layers; ++l)
      if (work[l] * ranks[k] > work[k] * ranks[l]) k = l;
    ++ranks[k], ++used;
  }
  int r = 0;
  for (int l = 0; l < n_layers; ++l) {
    P.first[l] = r, P.count[l] = ranks[l];
    P.ch_per_cta[l] = (C[l] + ranks[l] - 1) / ranks[l];
    for (int j = 0; j < ranks[l]; ++j) P.rank_layer[r++] = l;
  }
  P.n_layers = n_layers, P.B = B, P.out_h = out_h, P.out_w = out_w;
  P.rows_per_cta = (out_h + SC_CLUSTER - 1) / SC_CLUSTER;
  P.alpha = alpha, P.alpha_mode = alpha_mode, P.thresh = thresh, P.band = band;
  P.inv_layers = ((n_layers & (n_layers - 1)) == 0) ? 1.0f / (float)n_layers : 0.f;
  P.n_layers_f = (float)n_layers;
  P.cam_out = cam_out, P.mask_out = mask_out, P.near_count = near_count;
  P.vec_ok = ((out_w & 3) == 0) && (((uintptr_t)cam_out & 15) == 0) && (((uintptr_t)mask_out & 3) == 0);
  static int usable[64] = {};  // per device: 0 unknown, 1 clusters of 16 can be scheduled, -1 they cannot
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0;
  if (usable[dev_id] == 0) usable[dev_id] = sc_launch_t<WSDL_F32, 4, 4>(P, s, true) == 0 ? 1 : -1;
  if (usable[dev_id] < 0) return 1;
  if (dtype == WSDL_F32) return B <= 8 ? sc_launch_t<WSDL_F32, 4, 8>(P, s, false) : sc_launch_t<WSDL_F32, 4, 4>(P, s, false);
  if (dtype == WSDL_BF16) return sc_launch_t<WSDL_BF16, 8, 4>(P, s, false);
  return sc_launch_t<WSDL_F16, 8, 4>(P, s, false);
}

}  // namespace wsdl

using namespace wsdl;

extern "C" size_t wsdl_layercam_workspace_bytes(const int* C, const int* h, const int* w, int n_layers, int B,
                                                int dtype) {
  LcPlan pl;
  if (lc_plan(C, h, w, n_layers, B, dtype, &pl) != 0) return 0;
  return lc_workspace_bytes(pl);
}

extern "C" int wsdl_layercam_fused(const void* const* act, const void* const* grad, const int* C, const int* h,
                                   const int* w, int n_layers, int B, int dtype, int out_h, int out_w, float alpha,
                                   int alpha_mode, float thresh, float near_band, float* cam_out, uint8_t* mask_out,
                                   unsigned long long* near_count, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (!act || !grad || !workspace) return WSDL_E_NULL;
  if (out_h < 1 || out_w < 1) return WSDL_E_SHAPE;
  if (alpha_mode != 0 && alpha_mode != 1) return WSDL_E_ARG;
  if (B > 65535) return WSDL_E_SHAPE;
  LcPlan pl;
  int rc = lc_plan(C, h, w, n_layers, B, dtype, &pl);
  if (rc) return rc;
  if (workspace_bytes < lc_workspace_bytes(pl)) return WSDL_E_WORKSPACE;
  const int esize = dtype == WSDL_F32 ? 4 : 2;
  const size_t need_align = (size_t)pl.vec * esize;
  for (int l = 0; l < n_layers; ++l) {
    if (!act[l] || !grad[l]) return WSDL_E_NULL;
    if (((uintptr_t)act[l] % need_align) || ((uintptr_t)grad[l] % need_align)) return WSDL_E_ALIGN;
    pl.P.L[l].act = act[l];
    pl.P.L[l].grad = grad[l];
  }
  if (cam_out && ((uintptr_t)cam_out % 4)) return WSDL_E_ALIGN;
  if (near_count && ((uintptr_t)near_count % 8)) return WSDL_E_ALIGN;
  cudaStream_t s = (cudaStream_t)stream;
  if (cam_out || mask_out || near_count) {  // small batches: one cluster launch per call (no workspace, no memset)
    const int rc_small = sc_launch(act, grad, C, h, w, n_layers, B, dtype, out_h, out_w, alpha, alpha_mode, thresh, near_band,
                                   cam_out, mask_out, near_count, s);
    if (rc_small != 1) return rc_small;
  }
  uintptr_t ws = ((uintptr_t)workspace + 255) / 256 * 256;
  pl.P.ctrl = reinterpret_cast<unsigned*>(ws);
  pl.P.low = reinterpret_cast<float*>(ws) + pl.ctrl_u32;
  pl.P.partial = pl.P.low + pl.low_f32;
  pl.P.alpha = alpha;
  pl.P.alpha_mode = alpha_mode;

  cudaError_t e = cudaMemsetAsync(pl.P.ctrl, 0, pl.ctrl_u32 * 4, s);
  if (e != cudaSuccess) return (int)e;
  if (dtype == WSDL_F32)
    launch_channel_sum<WSDL_F32>(pl, s);
  else if (dtype == WSDL_BF16)
    launch_channel_sum<WSDL_BF16>(pl, s);
  else
    launch_channel_sum<WSDL_F16>(pl, s);
  WSDL_LAUNCH_CHECK();

  if (!cam_out && !mask_out && !near_count) return 0;
  UpParams U;
  for (int l = 0; l < n_layers; ++l) {
    U.L[l].low = pl.P.low + pl.P.L[l].low_off;
    U.L[l].h = h[l];
    U.L[l].w = w[l];
    U.L[l].scale_y = (float)h[l] / (float)out_h;
    U.L[l].scale_x = (float)w[l] / (float)out_w;
  }
  U.n_layers = n_layers;
  U.B = B;
  U.out_h = out_h;
  U.out_w = out_w;
  U.alpha = alpha;
  U.alpha_mode = alpha_mode;
  U.thresh = thresh;
  U.band = near_band;
  U.inv_layers = ((n_layers & (n_layers - 1)) == 0) ? 1.0f / (float)n_layers : 0.f;
  U.n_layers_f = (float)n_layers;
  U.cam_out = cam_out;
  U.mask_out = mask_out;
  U.near_count = near_count;
  U.vec_ok = ((out_w & 3) == 0) && (((uintptr_t)cam_out & 15) == 0) && (((uintptr_t)mask_out & 3) == 0);
  dim3 grid((out_w + UP_THREADS - 1) / UP_THREADS, (out_h + UP_ROWS - 1) / UP_ROWS, B);
  if (grid.y > 65535) return WSDL_E_SHAPE;
  dim3 grid4((out_w + UP_THREADS * UP_COLS - 1) / (UP_THREADS * UP_COLS), grid.y, B);  // 4 columns per thread
  switch (n_layers) {
    case 1: layercam_upsample_fuse<1><<<grid4, UP_THREADS, 0, s>>>(U); break;
    case 2: layercam_upsample_fuse<2><<<grid4, UP_THREADS, 0, s>>>(U); break;
    case 3: layercam_upsample_fuse<3><<<grid4, UP_THREADS, 0, s>>>(U); break;
    case 4: layercam_upsample_fuse<4><<<grid4, UP_THREADS, 0, s>>>(U); break;
    default: layercam_upsample_fuse_generic<<<grid, UP_THREADS, 0, s>>>(U); break;
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}

extern "C" int wsdl_threshold_mask(const float* cam, size_t n, float thresh, float near_band, uint8_t* mask_out,
                                   unsigned long long* near_count, void* stream) {
  if (!cam) return WSDL_E_NULL;
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)WSDL_NUM_SMS * 16) blocks = (size_t)WSDL_NUM_SMS * 16;
  threshold_mask_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(cam, n, thresh, near_band, mask_out,
                                                                           near_count);
  WSDL_LAUNCH_CHECK();
  return 0;
}
