// Weak-supervision loss of a two-class batch in ONE persistent launch, sm_100a:
//   cut loss on the logits        (LocalNormalizedCutLoss.forward, TraditionalModel/AlternatingDirectionCutLoss.py:71-105)
// + boundary loss per image on softmax(logits)
//                                 (ConstrainToBoundaryLossSingle.forward, AlternatingDirectionBoundaryLoss.py:20-70)
// + optionally cross-entropy against the pseudo-labels (the criterion of SegmentationModel.py:96-113),
// forward values and d(total)/d logits.  Same arithmetic as pairwise_dual_kernel (pairwise_sym.cu: pair symmetry,
// Euler loss, border multiplicities, sentinel colour, packed FADD2/FFMA2 march -- the device code is shared through
// pairwise_march.cuh).  What changes is WHEN things happen.
//
// Per-warp traces of the one-tile-per-CTA kernel (profiles/r01_f_trace_dual.txt) show a CTA marching for 6.6 us of
// a 14.6 us life: it waits for its tile (1.3-5.6 us), converts it (1.8 us), marches, then finishes segment heads,
// band columns and its loss partial (1.6-2 us) -- and all the time it holds a third of the SM's registers.  Here
//   * one CTA per SM lives for the whole launch: three MARCH GROUPS of four warps (the old CTA; 12 marching warps at
//     168 registers are all the registers an SM has, so there is no producer warp);
//   * tiles come from a global queue (short row blocks last); shared memory is FOUR tile slots for three groups.
//     f32 inputs: the slots rotate.  A group that finishes a tile hands its slot to the next tile of the queue (one
//     TMA box per raw plane, started at once) and takes the slot that was loading meanwhile; it converts that tile
//     in place (image * sqrt(-kc), logits -> p0) and marches.  Nobody ever waits for HBM after the first tile;
//   * bf16 logits / u8 images (the training step's and the dataset's own types): raw tiles do not have the f32
//     layout, so slot 3 is a staging buffer; the group that takes a staged tile converts it OUT OF PLACE into its
//     own slot, and the moment its four warps have read the staging buffer its last warp starts the TMA that
//     refills it.  Types are told apart in the conversion, nowhere else;
//   * the groups drift out of phase by construction: while one converts or finishes heads, the other two march,
//     and the latency-bound phases cost issue slots only;
//   * loss partials are per tile (bit-reproducible whichever group computed them); the last CTA to check in adds
//     them per image in a fixed order in double -- nobody spins on anybody.
#include <cuda.h>
#include <string.h>

#include "pairwise.cuh"
#include "pairwise_march.cuh"

namespace wsdl {

constexpr int SG_GROUPS = 3;                              // march groups per CTA
constexpr int SG_GTHREADS = PS_THREADS;                   // 128: 8 half-warp segments
constexpr int SG_THREADS = SG_GROUPS * SG_GTHREADS;
// "tile staged" barriers, one per ticket modulo the ring.  Ticket k + n_stages is staged by the group that consumed
// ticket k, so a barrier is handed to ticket k + 8 long after every waiter of ticket k has seen it complete: no parity
// aliasing with up to three groups waiting on consecutive tickets.
constexpr int SG_RING = 8;
constexpr int SG_MAX_STAGES = 2;                          // staging buffers (2 when the raw tile is small: u8 + bf16)
constexpr int SG_SLOTS = SG_GROUPS + 1;                   // tile slots: one per group + a spare / the staging area
constexpr int SG_SLOT_FLOATS = 5 * PS_PLANE;              // a raw f32 tile: 3 image planes + 2 logit planes
constexpr int SG_HEAD_FLOATS = (PS_SEGS - 1) * 2 * 2 * 64;
constexpr int SG_GBAND_FLOATS = PS_ROWS * 12;
static_assert(SG_HEAD_FLOATS + SG_GBAND_FLOATS <= PS_PLANE, "segment heads + band G live in the dead fifth plane");
constexpr int SG_WX_FLOATS = 2 * 6 * 10;
constexpr int SG_STAGE_BYTES = SG_SLOT_FLOATS * 4;
constexpr size_t SG_SMEM_BYTES = (size_t)SG_SLOTS * SG_SLOT_FLOATS * 4 + SG_WX_FLOATS * 4;
constexpr unsigned SG_SPIN_LIMIT = 400000u;  // x 20 us suspend hint: a barrier that never opens traps instead of hanging

struct SgParams {
  // inputs / outputs (device pointers)
  const void* labels;        // nullable: no cross-entropy term
  void* grad;                // nullable: forward only
  const float* go_cut;       // nullable upstream gradient of the cut loss (1 float)
  const float* go_bnd;       // nullable upstream gradients of the boundary losses (B floats)
  const float* ce_inv_n_dev; // nullable device scalar: 1 / (number of labels that are not ignore_index)
  float* loss_cut;           // 1
  float* loss_bnd;           // B
  float* loss_ce;            // 1, nullable
  float* loss_total;         // 1, nullable: lam_ce * ce + go_cut * cut + sum_b go_bnd[b] * bnd[b]
  float* partial;            // [n_tiles][3]: cut, boundary, cross-entropy sums of a tile
  unsigned* ctrl;            // [1] CTA check-in; 0 before and after a launch
  long long ignore_index;
  int B, H, W;
  int n_x, nb, n_tiles, S;   // column tiles per image, row blocks per column tile, tiles, rows per segment
  int logit_dtype, image_dtype, label_dtype, grad_dtype;  // WSDL_F32 / WSDL_BF16; WSDL_F32 / WSDL_U8; WSDL_U8 / WSDL_I64
  int img_pitch_b, img_plane_b, val_pitch_b, val_plane_b;  // staged raw layout, bytes
  int img_align, val_align;  // elements per 16 bytes: a TMA box must START on a 16-byte boundary of its row
  int stage_bytes, n_stages;
  unsigned stagger_ns;       // hold-back between the first tiles of a CTA's groups
  int rotate;                // 1: f32 logits and images -- tiles load straight into a spare slot and are converted in place
  unsigned tile_tx_bytes;
  float img_scale;           // sqrt(-kc_cut) (u8 images: applied after the exact / 255)
  float ratio, ksu_b, g1b, g4b, l1g_b, l32;
  float lam_ce, ce_inv_n_host;
  double kappa_cut, kappa_bnd;
};

#ifdef WSDL_PS_TRACE  // debug aid (scripts/trace_stream.py): per group and tile, timestamps of the phase boundaries
__device__ unsigned long long sg_trace_buf[WSDL_NUM_SMS * SG_GROUPS * 8 * 8];
#define SG_TR(slot_)                                                                                       \
  do {                                                                                                     \
    if (gt == 0 && n_done < 8) {                                                                           \
      unsigned long long t__;                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                              \
      sg_trace_buf[((blockIdx.x * SG_GROUPS + g) * 8 + n_done) * 8 + (slot_)] = t__;                       \
    }                                                                                                      \
  } while (0)
#else
#define SG_TR(slot_)
#endif

// Kernel variants.  The conversion loop is latency bound and the march loop issue bound: a run-time `if` on a type costs
// the first its load/compute overlap (the compiler cannot hoist loads across the branch) and the second ~180
// instructions per step, so the hot configurations are compiled with their types known:
//   0: f32 logits + f32 images, no labels (the cut + boundary launch of BASELINE configs[1])
//   1: f32 logits + f32 images + labels (cross-entropy fused in)
//   2: anything else (u8 images, ...): types read from the parameters
//   3: bf16 logits + f32 images + labels, bf16 gradient (the training step under bf16 autocast, BASELINE config 4)
template <int MODE>
struct SgMode {
  static constexpr bool known = MODE < 2;   // f32 raw tiles in the working layout: rotating slots, in-place conversion
  static constexpr bool typed = MODE != 2;  // element types known at compile time
  __device__ static __forceinline__ bool ce(const SgParams& P) { return MODE == 0 ? false : (typed ? true : P.labels != nullptr); }
  __device__ static __forceinline__ bool rotate(const SgParams& P) { return known ? true : (MODE == 3 ? false : P.rotate != 0); }
  __device__ static __forceinline__ bool img_f32(const SgParams& P) { return typed ? true : P.image_dtype == WSDL_F32; }
  __device__ static __forceinline__ bool val_f32(const SgParams& P) { return known ? true : (MODE == 3 ? false : P.logit_dtype == WSDL_F32); }
  __device__ static __forceinline__ bool grad_f32(const SgParams& P) { return known ? true : (MODE == 3 ? false : P.grad_dtype == WSDL_F32); }
};

__device__ __forceinline__ void sg_bar_wait(unsigned bar, unsigned parity) {
  unsigned spins = 0;
  unsigned ok;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok && ++spins > SG_SPIN_LIMIT) __trap();
  } while (!ok);
}

__device__ __forceinline__ void sg_group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + 4 * g) : "memory"); }

// First column of a tile's TMA box: x0 - 4 rounded down to the 16-byte grid of the element type (f32: x0 - 4 itself)
__device__ __forceinline__ int sg_box_x(int x0, int align) { return ((x0 - 4 + 64 * align) / align - 64) * align; }

__device__ __forceinline__ float sg_bf16_lo(unsigned w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float sg_bf16_hi(unsigned w) { return __uint_as_float(w & 0xffff0000u); }

// Labels of the 4 pixels (y, xs .. xs+3) of image b, one per byte: class ids 0 / 1; anything else (ignore_index, a
// position outside the image) -> 2 = "no cross-entropy here".  row0 = (b H + y) W.  xs is even and W a multiple of 4
// (the launcher checks), so the pixels come in two pairs that are inside or outside the image together: byte labels
// are two 16-bit loads.
__device__ __forceinline__ unsigned sg_label_class(long long v, long long ignore) {
  return (v == ignore || v < 0 || v > 1) ? 2u : (unsigned)v;
}
__device__ __forceinline__ unsigned sg_load_labels4(const SgParams& P, size_t row0, int xs) {
  unsigned out = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int x = xs + 2 * h;
    unsigned c0 = 2u, c1 = 2u;
    if (x >= 0 && x < P.W) {
      if (P.label_dtype == WSDL_U8) {
        const unsigned v = __ldg(reinterpret_cast<const unsigned short*>(reinterpret_cast<const unsigned char*>(P.labels) + row0 + x));
        c0 = sg_label_class((long long)(v & 0xffu), P.ignore_index), c1 = sg_label_class((long long)(v >> 8), P.ignore_index);
      } else {
        const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(reinterpret_cast<const long long*>(P.labels) + row0 + x));
        c0 = sg_label_class(v.x, P.ignore_index), c1 = sg_label_class(v.y, P.ignore_index);
      }
    }
    out |= (c0 | (c1 << 8)) << (16 * h);
  }
  return out;
}

// Rows [r0, r1) of the staged raw tile -> the group's planes: image * sqrt(-kc) (sentinel outside the image), logits -> p0.
// With labels: also the cross-entropy VALUE of the tile's own pixels, from z = v1 - v0 (not from the rounded p0).
template <int MODE>
__device__ __forceinline__ void sg_rows_convert(const SgParams& P, const PsBlk& K, const unsigned char* raw, float* s_img,
                                                float* s_p, int r0, int r1, int lane, float& lsum_ce) {
  using M = SgMode<MODE>;
  const int H = P.H, W = P.W;
  const float sc = P.img_scale;
  const bool want_ce = M::ce(P);
  const bool img_f32 = M::img_f32(P), val_f32 = M::val_f32(P);
  // the staged boxes start on 16-byte boundaries: canonical column 0 (image x0 - 4) sits `shift` elements into a raw row
  const int ib = M::known ? 0 : (K.x0 - 4 - sg_box_x(K.x0, P.img_align)) * (img_f32 ? 4 : 1);
  const int vb = M::known ? 0 : (K.x0 - 4 - sg_box_x(K.x0, P.val_align)) * (val_f32 ? 4 : 2);
  const int img_plane_b = M::known ? PS_PLANE * 4 : P.img_plane_b, val_plane_b = M::known ? PS_PLANE * 4 : P.val_plane_b;
  const int img_pitch_b = M::known ? PS_PITCH * 4 : P.img_pitch_b, val_pitch_b = M::known ? PS_PITCH * 4 : P.val_pitch_b;
#pragma unroll 4
  for (int it = r0 * PS_Q + lane; it < r1 * PS_Q; it += 32) {
    const int t = it / PS_Q, q = it - t * PS_Q;
    const int y = K.ys - 2 + t, xb = K.x0 - 4 + 4 * q;
    const bool row_in = y >= 0 && y < H;
    float4 v[3];
    if (img_f32) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float4 r = *reinterpret_cast<const float4*>(raw + c * img_plane_b + t * img_pitch_b + ib + 16 * q);
        v[c] = make_float4(r.x * sc, r.y * sc, r.z * sc, r.w * sc);
      }
    } else {  // u8: the dataset's 8-bit pixels; value / 255 exactly as ToTensor's .div(255), then the kernel scale
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const unsigned r = *reinterpret_cast<const unsigned*>(raw + c * img_plane_b + t * img_pitch_b + ib + 4 * q);
        v[c] = make_float4(__fdiv_rn((float)(r & 0xffu), 255.f) * sc, __fdiv_rn((float)((r >> 8) & 0xffu), 255.f) * sc,
                           __fdiv_rn((float)((r >> 16) & 0xffu), 255.f) * sc, __fdiv_rn((float)(r >> 24), 255.f) * sc);
      }
    }
    if (!row_in || xb < 0 || xb + 3 >= W) {
      const bool all_out = !row_in || xb + 3 < 0 || xb >= W;
      if (all_out) {
        v[0] = make_float4(PS_SENTINEL, PS_SENTINEL, PS_SENTINEL, PS_SENTINEL);
      } else {
        if (xb + 0 < 0 || xb + 0 >= W) v[0].x = PS_SENTINEL;
        if (xb + 1 < 0 || xb + 1 >= W) v[0].y = PS_SENTINEL;
        if (xb + 2 < 0 || xb + 2 >= W) v[0].z = PS_SENTINEL;
        if (xb + 3 < 0 || xb + 3 >= W) v[0].w = PS_SENTINEL;
      }
    }
    const int so = it * 4;
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(s_img + c * PS_PLANE + so) = v[c];
    float z[4];  // v1 - v0
    const unsigned char* rv = raw + 3 * img_plane_b + vb;
    if (val_f32) {
      const float4 a = *reinterpret_cast<const float4*>(rv + t * val_pitch_b + 16 * q);
      const float4 b = *reinterpret_cast<const float4*>(rv + val_plane_b + t * val_pitch_b + 16 * q);
      z[0] = b.x - a.x, z[1] = b.y - a.y, z[2] = b.z - a.z, z[3] = b.w - a.w;
    } else {
      const uint2 a = *reinterpret_cast<const uint2*>(rv + t * val_pitch_b + 8 * q);
      const uint2 b = *reinterpret_cast<const uint2*>(rv + val_plane_b + t * val_pitch_b + 8 * q);
      z[0] = sg_bf16_lo(b.x) - sg_bf16_lo(a.x), z[1] = sg_bf16_hi(b.x) - sg_bf16_hi(a.x);
      z[2] = sg_bf16_lo(b.y) - sg_bf16_lo(a.y), z[3] = sg_bf16_hi(b.y) - sg_bf16_hi(a.y);
    }
    float4 p0;  // p0 = 1 / (1 + e^(v1 - v0))
    p0.x = rcp_approx(1.f + ex2_approx(z[0] * LOG2E));
    p0.y = rcp_approx(1.f + ex2_approx(z[1] * LOG2E));
    p0.z = rcp_approx(1.f + ex2_approx(z[2] * LOG2E));
    p0.w = rcp_approx(1.f + ex2_approx(z[3] * LOG2E));
    *reinterpret_cast<float4*>(s_p + so) = p0;
    if (want_ce && q >= 1 && q <= 15 && t >= 2 && t < K.nc && xb < W) {  // a group of the tile's own pixels
      const unsigned lab = sg_load_labels4(P, ((size_t)K.b * H + y) * W, xb);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned c = (lab >> (8 * j)) & 0xffu;
        if (c < 2u) {  // -log p_c = softplus(+z) for class 0, softplus(-z) for class 1
          // log(1 + t), t = e^-|s| in (0, 1]: two MUFU; the absolute error (< 1e-7) is what matters in a mean of O(1) terms
          const float s = c == 0u ? z[j] : -z[j];
          lsum_ce += fmaxf(s, 0.f) + __logf(1.f + __expf(-fabsf(s)));
        }
      }
    }
  }
}

// Final value of one pixel's gradient w.r.t. logit 0 (logit 1 gets the negative): pairwise part + cross-entropy part.
__device__ __forceinline__ float sg_px_grad(bool want_ce, float p0, float gc, float gb, float scale_c, float scale_b,
                                            float cw, unsigned lab) {
  const float p1 = 1.f - p0;
  float o = 2.f * p0 * p1 * fmaf(scale_c, gc, scale_b * gb);
  if (want_ce && lab < 2u) o = fmaf(cw, lab == 0u ? -p1 : p0, o);  // cw (p0 - [label == 0])
  return o;
}

// 4 finished pixels of centre row t (columns xs .. xs+3 of image row y)
template <int MODE>
__device__ __forceinline__ void sg_emit(const SgParams& P, const PsBlk& K, float scale_b, float cw, int t, int strip,
                                        int okmask, unsigned labs, const float (&G)[4][2], const float (&pc)[4],
                                        float* s_gband, float& lsum_c, float& lsum_b) {
  const int H = P.H, W = P.W;
  const int y = K.ys - 2 + t;
  const int xs = K.x0 - 2 + 4 * strip;
  if (K.xband) {  // block-uniform: the band-column pass finishes these pixels from their G
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int slot = ps_band_slot(xs + j, W);
      if (((okmask >> j) & 1) && slot >= 0) {
        s_gband[t * 12 + slot * 2 + 0] = G[j][0];
        s_gband[t * 12 + slot * 2 + 1] = G[j][1];
      }
    }
  }
  float out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float p0 = pc[j];
    out[j] = sg_px_grad(SgMode<MODE>::ce(P), p0, G[j][0], G[j][1], K.scale2, scale_b, cw, (labs >> (8 * j)) & 0xffu);
    if ((okmask >> j) & 1) {
      lsum_c = fmaf(p0 + p0 - 1.f, G[j][0], lsum_c);
      lsum_b = fmaf(p0 + p0 - 1.f, G[j][1], lsum_b);
    }
  }
  if (!P.grad) return;
  const size_t plane = (size_t)H * W;
  const size_t o = (size_t)K.b * 2 * plane + (size_t)y * W + xs;  // xs is even and W % 4 == 0: pairs never straddle W
  if (SgMode<MODE>::grad_f32(P)) {
    float* go = reinterpret_cast<float*>(P.grad) + o;
    if (okmask & 1) {
      *reinterpret_cast<float2*>(go) = make_float2(out[0], out[1]);
      *reinterpret_cast<float2*>(go + plane) = make_float2(-out[0], -out[1]);
    }
    if (okmask & 4) {
      *reinterpret_cast<float2*>(go + 2) = make_float2(out[2], out[3]);
      *reinterpret_cast<float2*>(go + plane + 2) = make_float2(-out[2], -out[3]);
    }
  } else {
    __nv_bfloat16* go = reinterpret_cast<__nv_bfloat16*>(P.grad) + o;
    if (okmask & 1) {
      *reinterpret_cast<__nv_bfloat162*>(go) = __floats2bfloat162_rn(out[0], out[1]);
      *reinterpret_cast<__nv_bfloat162*>(go + plane) = __floats2bfloat162_rn(-out[0], -out[1]);
    }
    if (okmask & 4) {
      *reinterpret_cast<__nv_bfloat162*>(go + 2) = __floats2bfloat162_rn(out[2], out[3]);
      *reinterpret_cast<__nv_bfloat162*>(go + plane + 2) = __floats2bfloat162_rn(-out[2], -out[3]);
    }
  }
}

// Column tile of the j-th tile of (image b, row block rb): rotated by b + rb.  A CTA takes tiles CTA + n * grid; with a
// plain j the 148 SMs / 4 column tiles of a 224-wide image would hand some CTAs nothing but border-column tiles (which
// carry the band pass) and others none.
__device__ __forceinline__ int sg_tile_column(int j, int b, int rb, int n_x) { return (j + b + rb) % n_x; }

// Stage ticket k (one thread): draw a tile id from the global queue and start the TMA loads of its five raw planes into
// `dst_smem`, or -- queue empty / `end` -- publish an end marker.  The destination is free: it has never been used, or
// the caller's group has finished with it (rotating slots), or the caller has seen its "empty" barrier complete.
__device__ __forceinline__ void sg_stage_tile(const SgParams& P, const CUtensorMap& tm_img, const CUtensorMap& tm_val,
                                              unsigned char* dst_smem, unsigned long long* s_full, int* s_tile,
                                              unsigned* s_next, unsigned k, bool end) {
  const unsigned full = ps_smem_u32(&s_full[k % SG_RING]);
  // Tiles are dealt to the CTAs round robin (tile = CTA + n * grid, short row blocks last) and to a CTA's groups on
  // demand.  A global first-come queue was measured and is worse: a CTA prefetches one tile ahead, early CTAs drew up to
  // 8 tiles where 5.2 is the mean, and their third round set the launch time.
  const int id = end ? P.n_tiles : (int)(blockIdx.x + atomicAdd(s_next, 1u) * gridDim.x);
  if (id >= P.n_tiles) {
    s_tile[k % SG_RING] = -1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full) : "memory");
    return;
  }
  s_tile[k % SG_RING] = id;
  const int per_rb = P.n_x * P.B;
  const int rb = id / per_rb, rest = id - rb * per_rb;
  const int b = rest / P.n_x, tx = sg_tile_column(rest - b * P.n_x, b, rb, P.n_x);
  const int x0 = tx * PS_TW, ys = rb * (PS_SEGS * P.S - 2);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(P.tile_tx_bytes) : "memory");
  const unsigned dst = ps_smem_u32(dst_smem);
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
    const int pl = c < 3 ? b * 3 + c : b * 2 + (c - 3);
    const unsigned d = dst + (c < 3 ? c * P.img_plane_b : 3 * P.img_plane_b + (c - 3) * P.val_plane_b);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
        "l"(tm), "r"(sg_box_x(x0, c < 3 ? P.img_align : P.val_align)), "r"(ys - 2), "r"(pl), "r"(full)
        : "memory");
  }
}

template <int MODE>
__global__ void __launch_bounds__(SG_THREADS, 1)
    weak_loss_stream_kernel(const __grid_constant__ SgParams P, const __grid_constant__ CUtensorMap tm_img,
                            const __grid_constant__ CUtensorMap tm_val) {
  extern __shared__ __align__(128) unsigned char sg_smem[];
  __shared__ __align__(8) unsigned long long s_full[SG_RING];        // tile k staged (TMA bytes landed / end marker)
  __shared__ __align__(8) unsigned long long s_empty[SG_MAX_STAGES]; // staging buffer read by all 4 warps of its group
  __shared__ int s_tile[SG_RING];
  __shared__ int s_slot[SG_RING];             // rotating slots: where ticket k was loaded
  __shared__ unsigned s_cons;                 // next ticket of the consumer groups
  __shared__ unsigned s_prod;                 // rotating slots: next ticket to stage
  __shared__ unsigned s_next;                 // tiles this CTA has drawn
  __shared__ unsigned s_ticket[SG_GROUPS];
  __shared__ float s_red[SG_GROUPS][3][4];
  __shared__ int s_last;
  __shared__ double s_dred[SG_THREADS / 32];

  using M = SgMode<MODE>;
  const bool rotate = M::rotate(P);
  const int tid = threadIdx.x;
  const int H = P.H, W = P.W, S = P.S;
  float* s_wx = reinterpret_cast<float*>(sg_smem) + SG_SLOTS * SG_SLOT_FLOATS;  // [2][6][2][5]: column weights, cut then boundary
  if (tid == 0) {
    for (int i = 0; i < SG_RING; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ps_smem_u32(&s_full[i])));
    for (int i = 0; i < SG_MAX_STAGES; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(ps_smem_u32(&s_empty[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_cons = 0;
    s_prod = SG_SLOTS;
    s_next = 0;
  }
  if (tid >= 128 && tid < 248) {  // column weights of the band slots (depend on W only): gamma = 1 (cut), gamma_b (boundary)
    const int i = tid - 128, which = i / 60, e = i - which * 60, slot = e / 10, rem = e - slot * 10, j = rem % 5;
    const float g1 = which ? P.g1b : 1.f, g4 = which ? P.g4b : 1.f;
    const int x = slot < 3 ? slot : W - 6 + slot, xb = x + j - 2;
    float w = 0.f;
    if (xb >= 0 && xb < W) w = rem < 5 ? ps_w1d(x, xb, W, g1, g4) : ps_w1d(xb, x, W, g1, g4);
    s_wx[i] = w;
  }
  __syncthreads();

  if (tid == 0) {
    if (rotate) {  // every slot starts loading a tile
      for (int k = 0; k < SG_SLOTS; ++k) {
        s_slot[k] = k;
        if (k > 0 && P.stagger_ns) __nanosleep(P.stagger_ns);  // de-phase the groups: see sg_launch
        sg_stage_tile(P, tm_img, tm_val, sg_smem + (size_t)k * SG_STAGE_BYTES, s_full, s_tile, &s_next, (unsigned)k, false);
      }
    } else {
      for (int k = 0; k < P.n_stages; ++k)
        sg_stage_tile(P, tm_img, tm_val, sg_smem + (size_t)SG_GROUPS * SG_STAGE_BYTES + (size_t)k * P.stage_bytes, s_full,
                      s_tile, &s_next, (unsigned)k, false);
    }
  }
  {
    const int g = tid / SG_GTHREADS, gt = tid - g * SG_GTHREADS;
    const int lane = gt & 31, warp = gt >> 5;
    const int seg = warp * 2 + (lane >> 4), strip = lane & 15;
    unsigned char* const stage0 = sg_smem + (size_t)SG_GROUPS * SG_STAGE_BYTES;  // staging area (not rotating): slot 3
    const int per_rb = P.n_x * P.B;
    const float ksu = P.ksu_b, ratio = P.ratio;
    const bool want_ce = M::ce(P);
    const float ce_inv_n = P.ce_inv_n_dev ? __ldg(P.ce_inv_n_dev) : P.ce_inv_n_host;
    const float cw = want_ce ? P.lam_ce * ce_inv_n : 0.f;
    int n_done = 0;
    (void)n_done;
    for (;;) {
      SG_TR(0);
      if (gt == 0) s_ticket[g] = atomicAdd(&s_cons, 1u);
      sg_group_sync(g);  // also: everybody is done with the previous tile's planes, heads and s_red
      const unsigned k = s_ticket[g];
      sg_bar_wait(ps_smem_u32(&s_full[k % SG_RING]), (k / SG_RING) & 1);
      SG_TR(1);
      const int id = s_tile[k % SG_RING];
      if (id < 0) {  // queue empty.  Staging mode: pass the end marker on to whoever draws the next ticket of this stage
        if (!rotate && gt == 0) sg_stage_tile(P, tm_img, tm_val, stage0, s_full, s_tile, &s_next, k + (unsigned)P.n_stages, true);
        break;
      }
      const unsigned stage = k % (unsigned)P.n_stages;
      const int slot = rotate ? s_slot[k % SG_RING] : g;
      float* region = reinterpret_cast<float*>(sg_smem) + (size_t)slot * SG_SLOT_FLOATS;
      float* s_img = region;                       // [3][PS_ROWS][PS_PITCH] scaled for sigma_cut
      float* s_p = region + 3 * PS_PLANE;          // [PS_ROWS][PS_PITCH] p0
      float* s_head = region + 4 * PS_PLANE;       // [7][2][2][64]: G (cut, boundary) of the segment heads, in the plane
      float* s_gband = s_head + SG_HEAD_FLOATS;    // [row t][6 slots][cut, boundary]     the second logit leaves dead
      const unsigned char* raw = rotate ? reinterpret_cast<const unsigned char*>(region) : stage0 + stage * (unsigned)P.stage_bytes;

      PsBlk K;
      const int rb = id / per_rb, rest = id - rb * per_rb;
      K.b = rest / P.n_x;
      K.x0 = sg_tile_column(rest - K.b * P.n_x, K.b, rb, P.n_x) * PS_TW;
      K.ys = rb * (PS_SEGS * S - 2);
      K.n = min(PS_SEGS * S - 2, H - K.ys);
      K.nc = K.n + 2;
      const int xe = min(K.x0 + PS_TW, W);
      K.xband = (K.x0 <= 2) || (xe - 1 >= W - 3);
      const int rows = K.nc + 2;
      K.scale2 = (float)(4.0 * P.kappa_cut) * (P.go_cut ? __ldg(P.go_cut) : 1.f);
      const float scale_b = (float)(4.0 * P.kappa_bnd) * (P.go_bnd ? __ldg(P.go_bnd + K.b) : 1.f);
      int okmask = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = 4 * strip + j;
        if (col >= 2 && col < 2 + PS_TW && K.x0 - 2 + col < W) okmask |= 1 << j;
      }
      float lsum_c = 0.f, lsum_b = 0.f, lsum_ce = 0.f;

      // ---- convert this warp's rows (in place / out of the staging buffer, which is then refilled) ----
      {
        const int r0 = min(2 * warp * S, rows), r1 = warp == PS_WARPS - 1 ? rows : min(2 * (warp + 1) * S, rows);
        const int re = min(r0 + 2, r1);  // the two rows the warp above looks ahead into
        sg_rows_convert<MODE>(P, K, raw, s_img, s_p, r0, re, lane, lsum_ce);
        if (!rotate && warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(1 + 4 * g + warp) : "memory");  // pairs with warp - 1
        sg_rows_convert<MODE>(P, K, raw, s_img, s_p, re, r1, lane, lsum_ce);
        __syncwarp();
        if (rotate) {
          sg_group_sync(g);  // the second logit plane has been read by everybody: it now holds heads and band G
        } else {
          if (lane == 0) {
            const unsigned empty = ps_smem_u32(&s_empty[stage]);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty) : "memory");
            if (warp == PS_WARPS - 1) {  // converts two rows more than the others: usually the last one in
              sg_bar_wait(empty, (k / (unsigned)P.n_stages) & 1);  // all four warps have read the staging buffer
              sg_stage_tile(P, tm_img, tm_val, stage0 + stage * (unsigned)P.stage_bytes, s_full, s_tile,
                            &s_next, k + (unsigned)P.n_stages, false);
            }
          }
          __syncwarp();
          if (warp < PS_WARPS - 1) asm volatile("bar.sync %0, 64;" ::"r"(1 + 4 * g + warp + 1) : "memory");
        }
      }

      SG_TR(2);
      // ---- march ----
      const int t0 = seg * S, t1 = min(t0 + S, K.nc);
      float oy[4][2], oz[4][2];
#pragma unroll
      for (int q = 0; q < 4; ++q) oy[q][0] = oy[q][1] = oz[q][0] = oz[q][1] = 0.f;
      if (warp * 2 * S < K.nc) {
        float2 A[4][2], Bq[4][2], Cq[4][2];
#pragma unroll
        for (int w = 0; w < 4; ++w)
#pragma unroll
          for (int c = 0; c < 2; ++c) A[w][c] = Bq[w][c] = Cq[w][c] = make_float2(0.f, 0.f);
        const int xs = K.x0 - 2 + 4 * strip;
#pragma unroll 1
        for (int s = 0; s < S; ++s) {
          const int t = t0 + s;
          const bool act = t < t1;
          float pc[4], own[4][2];
          unsigned labs = 0x02020202u;
          if (act) {
            const int y = K.ys - 2 + t;
            // labels of the row this step finishes: requested now, used ~700 instructions later
            if (want_ce && s >= 2 && okmask) labs = sg_load_labels4(P, ((size_t)K.b * H + y) * W, xs);
            const bool r1 = (y == 1 || y == H - 2);
            const float l0c = r1 ? 1.f : 0.f, l0b = r1 ? P.l1g_b : 0.f;  // log2(1 + gamma^4): gamma = 1 for the cut loss
            const float l1 = (y == 0 || y == H - 2) ? P.l32 : 0.f;
            const float l2 = (y == 0 || y == H - 3) ? P.l32 : 0.f;
            PsKsDual ks;
            ks.ca = l0c, ks.cb = l1, ks.cc = l2;
            const float ra = ratio * l0c, rb2 = ratio * l1, rc = ratio * l2;
            ks.a1 = ksu + l0b - ra, ks.a4 = 4.f * ksu + l0b - ra;
            ks.b0 = ksu + l1 - rb2, ks.b1 = 2.f * ksu + l1 - rb2, ks.b4 = 5.f * ksu + l1 - rb2;
            ks.c0 = 4.f * ksu + l2 - rc, ks.c1 = 5.f * ksu + l2 - rc, ks.c4 = 8.f * ksu + l2 - rc;
            ps_step_dual2(A, Bq, Cq, pc, s_img, s_p, t * PS_PITCH + 4 * strip, ks, ratio);
          }
          ps_exchangep<2>(A, own, strip);
          if (act) {
            if (s >= 2) {
              sg_emit<MODE>(P, K, scale_b, cw, t, strip, okmask, labs, own, pc, s_gband, lsum_c, lsum_b);
            } else if (seg > 0) {
#pragma unroll
              for (int c = 0; c < 2; ++c)
                *reinterpret_cast<float4*>(s_head + (((seg - 1) * 2 + s) * 2 + c) * 64 + 4 * strip) =
                    make_float4(own[0][c], own[1][c], own[2][c], own[3][c]);
            }
          }
#pragma unroll
          for (int w = 0; w < 4; ++w)
#pragma unroll
            for (int c = 0; c < 2; ++c) A[w][c] = Bq[w][c], Bq[w][c] = Cq[w][c], Cq[w][c] = make_float2(0.f, 0.f);
        }
        ps_exchangep<2>(A, oy, strip);
        ps_exchangep<2>(Bq, oz, strip);
      }
      SG_TR(3);
      sg_group_sync(g);  // every head row holds its own segment's part
      SG_TR(4);
      if (seg < PS_SEGS - 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (t0 + S < K.nc) {
            float4* h = reinterpret_cast<float4*>(s_head + ((seg * 2 + 0) * 2 + c) * 64 + 4 * strip);
            float4 v = *h;
            v.x += oy[0][c], v.y += oy[1][c], v.z += oy[2][c], v.w += oy[3][c];
            *h = v;
          }
          if (t0 + S + 1 < K.nc) {
            float4* h = reinterpret_cast<float4*>(s_head + ((seg * 2 + 1) * 2 + c) * 64 + 4 * strip);
            float4 v = *h;
            v.x += oz[0][c], v.y += oz[1][c], v.z += oz[2][c], v.w += oz[3][c];
            *h = v;
          }
        }
      }
      sg_group_sync(g);

      // ---- the first two rows of segments 1..7, now complete ----
      if (seg > 0) {
        const int xs = K.x0 - 2 + 4 * strip;
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int t = t0 + i;
          if (t < t1) {
            float G[4][2], pc[4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const float4 h = *reinterpret_cast<const float4*>(s_head + (((seg - 1) * 2 + i) * 2 + c) * 64 + 4 * strip);
              G[0][c] = h.x, G[1][c] = h.y, G[2][c] = h.z, G[3][c] = h.w;
            }
            const float2 p01 = *reinterpret_cast<const float2*>(s_p + t * PS_PITCH + 4 * strip + 2);
            const float2 p23 = *reinterpret_cast<const float2*>(s_p + t * PS_PITCH + 4 * strip + 4);
            pc[0] = p01.x, pc[1] = p01.y, pc[2] = p23.x, pc[3] = p23.y;
            unsigned labs = 0x02020202u;
            if (want_ce && okmask) labs = sg_load_labels4(P, ((size_t)K.b * H + (K.ys - 2 + t)) * W, xs);
            sg_emit<MODE>(P, K, scale_b, cw, t, strip, okmask, labs, G, pc, s_gband, lsum_c, lsum_b);
          }
        }
      }

      SG_TR(5);
      // ---- band columns (and corners) ----
      if (K.xband) {
        sg_group_sync(g);
        const int nlo = max(0, min(3, xe) - K.x0), hi0 = max(W - 3, K.x0), nhi = max(0, xe - hi0);
        const int ncb = nlo + nhi;
        const size_t plane = (size_t)H * W;
        for (int i = gt; i < ncb * K.n; i += SG_GTHREADS) {
          const int ty = i / ncb, kk = i - ty * ncb;
          const int x = kk < nlo ? K.x0 + kk : hi0 + (kk - nlo), y = K.ys + ty;
          const int slot = ps_band_slot(x, W);
          float ac, ab, p0;  // both losses from one set of partner loads and one squared distance
          ps_xfix_dual(H, P.g1b, P.g4b, P.ratio, s_img, s_p, s_wx + slot * 10, s_wx + 60 + slot * 10, K.ys, K.x0, y, x, ac, ab, p0);
          lsum_c = fmaf(p0 + p0 - 1.f, ac, lsum_c);  // the losses are linear in G: the corrections' share
          lsum_b = fmaf(p0 + p0 - 1.f, ab, lsum_b);
          if (P.grad) {
            const float gc = ac + s_gband[(ty + 2) * 12 + slot * 2 + 0];
            const float gb = ab + s_gband[(ty + 2) * 12 + slot * 2 + 1];
            unsigned lab = 2u;
            if (want_ce) {
              const size_t pix = ((size_t)K.b * H + y) * W + x;
              lab = sg_label_class(P.label_dtype == WSDL_U8 ? (long long)__ldg(reinterpret_cast<const unsigned char*>(P.labels) + pix)
                                                            : __ldg(reinterpret_cast<const long long*>(P.labels) + pix),
                                   P.ignore_index);
            }
            const float o = sg_px_grad(want_ce, p0, gc, gb, K.scale2, scale_b, cw, lab);
            const size_t off = (size_t)K.b * 2 * plane + (size_t)y * W + x;
            if (M::grad_f32(P)) {
              float* go = reinterpret_cast<float*>(P.grad) + off;
              go[0] = o, go[plane] = -o;
            } else {
              __nv_bfloat16* go = reinterpret_cast<__nv_bfloat16*>(P.grad) + off;
              go[0] = __float2bfloat16_rn(o), go[plane] = __float2bfloat16_rn(-o);
            }
          }
        }
      }

      SG_TR(6);
      // ---- the tile's three sums ----
      {
        const float wc = warp_sum(lsum_c), wb = warp_sum(lsum_b), we = warp_sum(lsum_ce);
        if (lane == 0) s_red[g][0][warp] = wc, s_red[g][1][warp] = wb, s_red[g][2][warp] = we;
        sg_group_sync(g);
        if (gt == 0) {
          float tc = 0.f, tb = 0.f, te = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) tc += s_red[g][0][i], tb += s_red[g][1][i], te += s_red[g][2][i];
          float* dst = P.partial + (size_t)id * 3;
          __stcg(dst + 0, 2.f * tc);
          __stcg(dst + 1, 2.f * tb);
          __stcg(dst + 2, te);
          if (rotate) {  // the slot is free (every thread of the group is past its last read: the barrier above): hand it
                         // to the next tile of the queue, which loads while the three groups work on theirs
            const unsigned n = atomicAdd(&s_prod, 1u);
            s_slot[n % SG_RING] = slot;
            sg_stage_tile(P, tm_img, tm_val, reinterpret_cast<unsigned char*>(region), s_full, s_tile, &s_next, n, false);
          }
        }
      }
      SG_TR(7);
      ++n_done;
    }
  }

  // ------------------------------------------------------------------ check in; the last CTA adds the partials up
  if ((tid & (SG_GTHREADS - 1)) == 0) __threadfence();  // this group leader's partial stores, before the CTA checks in
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned t = atomicAdd(P.ctrl + 1, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) P.ctrl[1] = 0u;  // every CTA has checked in: ready for the next launch
  const int warp = tid >> 5, lane = tid & 31, n_warps = SG_THREADS / 32;
  const int per_rb = P.n_x * P.B;
  double wtot_c = 0.0, wtot_e = 0.0, wtot_b = 0.0;
  for (int b = warp; b < P.B; b += n_warps) {  // image by image, tiles in a fixed order
    double ac = 0.0, ab = 0.0, ae = 0.0;
    const int kpi = P.nb * P.n_x;
    for (int i = lane; i < kpi; i += 32) {
      const int rb = i / P.n_x, tx = i - rb * P.n_x;
      const float* src = P.partial + ((size_t)rb * per_rb + (size_t)b * P.n_x + tx) * 3;
      ac += (double)ld_cg_f32(src + 0);
      ab += (double)ld_cg_f32(src + 1);
      ae += (double)ld_cg_f32(src + 2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ac += __shfl_xor_sync(0xffffffffu, ac, o);
      ab += __shfl_xor_sync(0xffffffffu, ab, o);
      ae += __shfl_xor_sync(0xffffffffu, ae, o);
    }
    const float lb = (float)(ab * P.kappa_bnd);
    if (lane == 0) P.loss_bnd[b] = lb;
    wtot_c += ac;
    wtot_e += ae;
    wtot_b += (double)lb * (double)(P.go_bnd ? __ldg(P.go_bnd + b) : 1.f);
  }
  // images were dealt to warps round-robin: add the warps' sums in warp order (fixed for a given B)
  __shared__ double s_fin[3];
  for (int what = 0; what < 3; ++what) {
    if (lane == 0) s_dred[warp] = what == 0 ? wtot_c : (what == 1 ? wtot_e : wtot_b);
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int i = 0; i < n_warps; ++i) t += s_dred[i];
      s_fin[what] = t;
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float lc = (float)(s_fin[0] * P.kappa_cut);
    P.loss_cut[0] = lc;
    float lce = 0.f;
    if (P.loss_ce) {
      const float inv_n = P.ce_inv_n_dev ? __ldg(P.ce_inv_n_dev) : P.ce_inv_n_host;
      lce = (float)(s_fin[1] * (double)inv_n);
      P.loss_ce[0] = lce;
    }
    if (P.loss_total)
      P.loss_total[0] = (float)((double)P.lam_ce * (double)lce + (double)(P.go_cut ? __ldg(P.go_cut) : 1.f) * (double)lc + s_fin[2]);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*SgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SgEncodeFn sg_encoder() {  // cuTensorMapEncodeTiled through the runtime: no link-time dependency on libcuda
  static const SgEncodeFn fn = []() -> SgEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (SgEncodeFn)p;
  }();
  return fn;
}

// (W, H, planes) tensor of `esize`-byte elements, box = box_w columns x rows x 1 plane, zero fill outside
static bool sg_encode(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int esize, int W, int H, long long planes,
                      int box_w, int rows) {
  const SgEncodeFn enc = sg_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * esize, (cuuint64_t)W * H * esize};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, dt, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Rows per segment: fewest modelled rounds x (march steps + fixed cost of a tile, in steps) over the 3 x 148 groups.
static int sg_rows_per_segment(int B, int H, int W) {
  static const int forced = WSDL_TUNE_INT("WSDL_SG_S", 0);
  if (forced >= 2 && forced <= PS_SMAX) return forced;
  const int n_x = (W + PS_TW - 1) / PS_TW;
  int best = PS_SMAX;
  double best_cost = 1e300;
  for (int S = PS_SMAX; S >= 2; --S) {
    const int nb = (H + PS_SEGS * S - 3) / (PS_SEGS * S - 2);
    const double tiles = (double)B * n_x * nb;
    const double groups = (double)SG_GROUPS * WSDL_NUM_SMS;
    const double rounds = tiles <= groups ? 1.0 : tiles / groups + 0.5;  // the queue balances; half a tile of ragged end
    const double cost = rounds * (S + 2.0);
    if (cost < best_cost - 1e-9) best_cost = cost, best = S;
  }
  return best;
}

size_t sg_workspace_floats(int B, int H, int W) {
  const int n_x = (W + PS_TW - 1) / PS_TW;
  const int nb_max = (H + PS_SEGS * 2 - 3) / (PS_SEGS * 2 - 2);  // S = 2 gives the most tiles
  return (size_t)3 * B * n_x * nb_max;
}

// 1: not this kernel's shape (the caller falls back), 0: launched, else a CUDA error
int sg_launch(const SgLaunch& L, cudaStream_t s) {
  if (L.H < 6 || L.W < 6 || L.B < 1) return 1;
  const int esz_img = L.image_dtype == WSDL_U8 ? 1 : 4, esz_val = L.logit_dtype == WSDL_BF16 ? 2 : 4;
  // TMA: 16-byte aligned bases and row strides
  if (((uintptr_t)L.images & 15) || ((uintptr_t)L.logits & 15) || ((L.W * esz_img) & 15) || ((L.W * esz_val) & 15)) return 1;
  if (L.labels && (L.W & 3)) return 1;
  if (L.grad && ((L.W & 3) || ((uintptr_t)L.grad & 7))) return 1;
  SgParams P;
  memset(&P, 0, sizeof(P));
  P.labels = L.labels, P.grad = L.grad, P.go_cut = L.go_cut, P.go_bnd = L.go_bnd, P.ce_inv_n_dev = L.ce_inv_n_dev;
  P.loss_cut = L.loss_cut, P.loss_bnd = L.loss_bnd, P.loss_ce = L.labels ? L.loss_ce : nullptr;
  P.loss_total = L.loss_total;
  P.partial = L.partial, P.ctrl = L.ctrl;
  P.ignore_index = L.ignore_index;
  P.B = L.B, P.H = L.H, P.W = L.W;
  P.n_x = (L.W + PS_TW - 1) / PS_TW;
  P.S = sg_rows_per_segment(L.B, L.H, L.W);
  P.nb = (L.H + PS_SEGS * P.S - 3) / (PS_SEGS * P.S - 2);
  const long long tiles = (long long)P.nb * P.n_x * L.B;
  if (tiles > 0x3fffffffLL) return 1;
  P.n_tiles = (int)tiles;
  P.logit_dtype = L.logit_dtype, P.image_dtype = L.image_dtype, P.label_dtype = L.label_dtype, P.grad_dtype = L.grad_dtype;
  const int rows = PS_SEGS * P.S + 2;
  // box = the 68 canonical columns + what it takes to start on a 16-byte boundary (u8: up to 12, bf16: up to 4 columns)
  const int img_box_w = L.image_dtype == WSDL_U8 ? 80 : PS_PITCH;
  const int val_box_w = L.logit_dtype == WSDL_BF16 ? 72 : PS_PITCH;
  P.img_pitch_b = img_box_w * esz_img, P.val_pitch_b = val_box_w * esz_val;
  P.img_align = 16 / esz_img, P.val_align = 16 / esz_val;
  P.rotate = (L.image_dtype == WSDL_F32 && L.logit_dtype == WSDL_F32 && !WSDL_TUNE_INT("WSDL_SG_NO_ROTATE", 0)) ? 1 : 0;
  if (P.rotate) {  // the raw tile IS the working layout: planes at the canonical offsets
    P.img_plane_b = P.val_plane_b = PS_PLANE * 4;
  } else {
    P.img_plane_b = (rows * P.img_pitch_b + 127) / 128 * 128;
    P.val_plane_b = (rows * P.val_pitch_b + 127) / 128 * 128;
  }
  P.stagger_ns = (unsigned)WSDL_TUNE_INT("WSDL_SG_STAGGER_NS", 0);
  P.stage_bytes = 3 * P.img_plane_b + 2 * P.val_plane_b;
  if (P.stage_bytes > SG_STAGE_BYTES) return 1;
  P.n_stages = SG_STAGE_BYTES / P.stage_bytes >= 2 ? 2 : 1;
  P.tile_tx_bytes = (unsigned)(rows * (3 * P.img_pitch_b + 2 * P.val_pitch_b));
  const float kc = -LOG2E / (2.f * L.sigma_cut * L.sigma_cut);
  P.img_scale = sqrtf(-kc);
  P.ratio = (L.sigma_cut * L.sigma_cut) / (L.sigma_bnd * L.sigma_bnd);
  const float inv_2ss = L.sigma_space > 0.f ? 1.f / (2.f * L.sigma_space * L.sigma_space) : 0.f;
  P.ksu_b = -LOG2E * inv_2ss;
  P.g1b = expf(-inv_2ss), P.g4b = expf(-4.f * inv_2ss);
  P.l1g_b = log2f(1.f + P.g4b);
  P.l32 = log2f(1.5f);
  P.lam_ce = L.lam_ce;
  P.ce_inv_n_host = 1.f / ((float)L.B * (float)L.H * (float)L.W);
  P.kappa_cut = 1.0 / (24.0 * (double)L.B * (double)L.H * (double)L.W * 2.0);
  P.kappa_bnd = 1.0 / (24.0 * (double)L.H * (double)L.W);
  CUtensorMap tm_img, tm_val;
  memset(&tm_img, 0, sizeof(tm_img)), memset(&tm_val, 0, sizeof(tm_val));
  if (!sg_encode(&tm_img, L.images, L.image_dtype == WSDL_U8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                 esz_img, L.W, L.H, 3LL * L.B, img_box_w, rows) ||
      !sg_encode(&tm_val, L.logits, L.logit_dtype == WSDL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                 esz_val, L.W, L.H, 2LL * L.B, val_box_w, rows))
    return 1;
  static bool attr_done[64] = {};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, attr_done[0] = false;
  if (!attr_done[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(weak_loss_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(weak_loss_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(weak_loss_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(weak_loss_stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_done[dev_id] = true;
  }
  int grid = (P.n_tiles + SG_GROUPS - 1) / SG_GROUPS;
  if (grid > WSDL_NUM_SMS) grid = WSDL_NUM_SMS;
  const bool f32_in = L.image_dtype == WSDL_F32 && L.logit_dtype == WSDL_F32 && (!L.grad || L.grad_dtype == WSDL_F32) && P.rotate;
  if (f32_in && !L.labels)
    weak_loss_stream_kernel<0><<<grid, SG_THREADS, SG_SMEM_BYTES, s>>>(P, tm_img, tm_val);
  else if (f32_in)
    weak_loss_stream_kernel<1><<<grid, SG_THREADS, SG_SMEM_BYTES, s>>>(P, tm_img, tm_val);
  else if (L.labels && L.image_dtype == WSDL_F32 && L.logit_dtype == WSDL_BF16 && (!L.grad || L.grad_dtype == WSDL_BF16))
    weak_loss_stream_kernel<3><<<grid, SG_THREADS, SG_SMEM_BYTES, s>>>(P, tm_img, tm_val);
  else
    weak_loss_stream_kernel<2><<<grid, SG_THREADS, SG_SMEM_BYTES, s>>>(P, tm_img, tm_val);
  WSDL_LAUNCH_CHECK();
  return 0;
}

}  // namespace wsdl

#ifdef WSDL_PS_TRACE
extern "C" int wsdl_sg_trace_read(unsigned long long* host, int n_words) {
  return (int)cudaMemcpyFromSymbol(host, wsdl::sg_trace_buf, (size_t)n_words * sizeof(unsigned long long));
}
#endif
