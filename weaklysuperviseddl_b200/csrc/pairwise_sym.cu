// Pair-symmetric cut / boundary loss (window 5, C <= 2), forward + backward in one launch, sm_100a.
//
// Same contract as pairwise.cu (reference: LocalNormalizedCutLoss.forward, TraditionalModel/
// AlternatingDirectionCutLoss.py:71-105; ConstrainToBoundaryLossSingle.forward,
// AlternatingDirectionBoundaryLoss.py:20-70) but organised around the instruction count, because on B200 the
// stencil is bound by the FP32/MUFU issue rate long before HBM (ncu: 24 exp + ~340 FP32 ops per pixel):
//
//  * k(z,y) is symmetric, so every unordered pixel pair is evaluated ONCE (12 forward offsets instead of 24 taps):
//    t = k s (p(z) - p(y)) is added to G(z) and subtracted from G(y).
//  * The loss is not accumulated per pair: L is a quadratic form in p, so  L = 1/2 sum_z p(z) . dL/dp(z)
//    (Euler); the kernel forms it from the finished gradient (with p shifted by 1/2: sum_z dL/dp(z) = 0).
//  * The image is pre-scaled by sqrt(-kc) in shared memory, so a pair costs 3 FADD + 3 FFMA + 1 MUFU.EX2 for k and
//    3 C more for the two scatters.
//  * Two classes behind a softmax (the reference's binary segmentation) satisfy p1 = 1 - p0, so G1 = -G0: only
//    channel 0 is accumulated (CS = 1 "stored channel"), and dL/dlogit0 = -dL/dlogit1 = 4 kappa p0 p1 G0.
//  * Two pairs with the same offset are evaluated per instruction with the two-lane FP32 forms of sm_100a (FADD2 /
//    FFMA2: the scalar lane rate at half the issue slots); see "packed (f32x2) form" below for the register layout.
//
// Work mapping.  A thread owns a strip of 4 columns and marches down S consecutive rows (a "segment"); the
// contributions it makes to rows t+1, t+2 live in three rotating accumulator rows in registers, the ones it makes
// to the 2 columns either side of its strip go to the neighbouring lanes by warp shuffle when a row completes.
// 16 strips (a half warp) span the 64 staged columns of a tile (60 owned + 2 halo each side), 8 segments span the
// block's rows; the first two rows of a segment are completed by what its upper neighbour adds to them, through
// shared memory after the march.  A CTA is one block: up to PS_CAP = 38 rows (fused kernel: 46) of one 60-column tile of one image (grid = row blocks x
// column tiles x images, a few CTAs per SM slot so that the loads of one overlap the march of the others); it
// pays 2 warm-up rows instead of a halo in y.
//
// Staging.  One thread issues a TMA tile load per input plane (cp.async.bulk.tensor.3d, box = 68 columns x all rows
// of the block, out-of-image positions zero-filled by the hardware) and every warp sleeps on the mbarrier.  Each warp
// then converts ITS rows in place (image * sqrt(-kc), logits -> probabilities, sentinel colour outside the image),
// hands its first two rows to the warp above through a named barrier, and starts marching: there is no CTA-wide
// barrier between the load and the march.  Inputs that TMA cannot address (W % 4 != 0, unaligned base) are loaded
// by the warps themselves, element by element, into the same layout.
//
// Borders.  Reflect padding only changes how often a pair of REAL pixels is counted: grouping the reference's
// sum over (centre, offset) by the pixel the reflected offset lands on gives
//     L = kappa sum_{a,b} Wy(ya->yb) Wx(xa->xb) kc(a,b) |p(a)-p(b)|^2,   W(u->v) = sum_{d=-2..2} [r(u+d) == v] gamma^(d^2),
// kc the colour-only affinity and gamma = exp(-1/(2 sigma_space^2)) (1 for the cut loss).  Away from the borders the
// unordered pair weight Wy Wx + Wy' Wx' is the interior 2 gamma^|d|^2; within 3 px of a border it is a multiple of it:
//   * rows: pairs {0,1}, {0,2} (and {H-1,H-2}, {H-1,H-3}) count 3/2, pairs inside row 1 (H-2) count 1+gamma^4.  The
//     march adds log2 of that factor to the exponent of k, per row step -- no extra instruction per pair;
//   * columns (and the corners, where the weight is not a product): after the march, a short pass over the band
//     pixels of a border tile adds (true weight - weight the march applied) kc (p(a)-p(b)) for the <= 24 partners of
//     each to the stored gradient (the gradient and the loss are linear in G).
// Positions outside the image carry a sentinel colour, so k underflows to exactly 0 for them.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "pairwise.cuh"
#include "pairwise_march.cuh"

#ifndef WSDL_X_SKIP
#define WSDL_X_SKIP 0  // measurement builds only (fused kernel, profiles/r02_g_phase_skip.txt): 1 no conversion, 2 no head rows, 4 no march, 8 no band pass
#endif

namespace wsdl {


#ifdef WSDL_PS_TRACE  // debug aid (scripts/trace_ctas.py): per-warp timestamps of the phase boundaries
__device__ unsigned long long ps_trace_buf[4096 * 4 * 16];
#define PS_TR(slot)                                                                        \
  do {                                                                                     \
    const unsigned cta__ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); \
    if ((threadIdx.x & 31) == 0 && cta__ < 4096) {                                         \
      unsigned long long t__;                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                              \
      ps_trace_buf[(cta__ * 4 + (threadIdx.x >> 5)) * 16 + (slot)] = t__;                  \
    }                                                                                      \
  } while (0)
#else
#define PS_TR(slot)
#endif

// One CTA = one block: rows [ys, ys + n) of one 60-column tile of one image, n <= PS_CAP.
template <int C, bool SOFTMAX>
__global__ void __launch_bounds__(PS_THREADS, PsCfg<C, SOFTMAX>::CTAS)
    pairwise_sym_kernel(const __grid_constant__ PsParams Q, const __grid_constant__ CUtensorMap tm_img,
                        const __grid_constant__ CUtensorMap tm_val) {
  constexpr int CS = PsCfg<C, SOFTMAX>::CS;
  extern __shared__ __align__(128) float ps_smem[];
  float* s_img = ps_smem;                                 // [3][PS_ROWS][PS_PITCH]: raw, then pre-scaled
  float* s_p = s_img + 3 * PS_PLANE;                      // [C][PS_ROWS][PS_PITCH]: raw values, then probabilities
  float* s_head = s_p + C * PS_PLANE;                     // [7][2][CS][64]: G of the first two rows of segments 1..7
  float* s_gband = s_head + (PS_SEGS - 1) * 2 * CS * 64;  // [6][CS][PS_CAP]: G of the band-column pixels
  float* s_wx = s_gband + 6 * CS * PS_CAP;                // H >= 10: the band tables fx1 | wy1 (268 floats); else [6][2][5]
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ float s_red[PS_THREADS / 32];
  __shared__ double s_dred[PS_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = warp * 2 + (lane >> 4), strip = lane & 15;
  const int H = Q.p.H, W = Q.p.W;
  const float ksu = Q.p.ks_unit;
  const int S = Q.S;
  PS_TR(0);

  PsBlk K;
  K.b = blockIdx.z;
  K.x0 = blockIdx.y * PS_TW;
  // full blocks of 8 S - 2 rows (every segment row is used), the remainder in the last block of the column
  K.ys = blockIdx.x * (PS_SEGS * S - 2);
  K.n = min(PS_SEGS * S - 2, H - K.ys);
  K.nc = K.n + 2;
  const int xe = min(K.x0 + PS_TW, W);  // owned pixels [x0, xe) x [ys, ys + n)
  K.xband = (K.x0 <= 2) || (xe - 1 >= W - 3);
  const int rows = K.nc + 2;  // staged rows in use: ys-2 .. ys+n+1

  // ---- tile load: one TMA box per plane (all 8S+2 rows of the launch's block shape), completion on s_bar ----
  if (Q.use_tma && tid == 0) {
    // The first wave's loads all queue on HBM at once; issued together they would all complete together (after the
    // whole wave's bytes have moved) with the SMs idle meanwhile.  CTAs are dispatched breadth first, so the k-th
    // CTA of an SM holds back k quarter-wave transfer times: the quarters complete in turn, the early ones march
    // under the later ones' traffic, and the SM's CTAs stay out of phase from then on.
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const unsigned bar = ps_smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_wait();  // programmatic dependent launch: nothing before this line touches global memory
    if (cta >= WSDL_NUM_SMS && cta < (unsigned)WSDL_NUM_SMS * PsCfg<C, SOFTMAX>::CTAS && Q.stagger_ns > 0)
      __nanosleep((cta / WSDL_NUM_SMS) * Q.stagger_ns);
    const unsigned bytes = (unsigned)((3 + C) * (PS_SEGS * S + 2) * PS_PITCH * sizeof(float));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
    for (int c = 0; c < 3 + C; ++c) {
      const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
      const int pl = c < 3 ? K.b * 3 + c : K.b * C + (c - 3);
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              ps_smem_u32(s_img + c * PS_PLANE)),
          "l"(tm), "r"(K.x0 - 4), "r"(K.ys - 2), "r"(pl), "r"(bar)
          : "memory");
    }
  }
  int okmask = 0;  // which of this thread's 4 columns are owned pixels of the image
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 4 * strip + j;  // staged column: 0,1 and 62,63 are halo
    if (col >= 2 && col < 2 + PS_TW && K.x0 - 2 + col < W) okmask |= 1 << j;
  }
  float lsum = 0.f;
  if (K.xband && H >= 10) {  // the band pass's weight tables, built on the host once per launch
    for (int i = tid; i < 48 + 220; i += PS_THREADS) s_wx[i] = i < 48 ? Q.fx1[i] : Q.wy1[i - 48];
  } else if (K.xband && tid >= 64 && tid < 124) {  // tiny images: column weights Wx(x -> x + j - 2), Wx(x + j - 2 -> x)
    const int e = tid - 64, slot = e / 10, rem = e - slot * 10, j = rem % 5;
    const int x = slot < 3 ? slot : W - 6 + slot, xb = x + j - 2;
    float w = 0.f;
    if (xb >= 0 && xb < W) w = rem < 5 ? ps_w1d(x, xb, W, Q.g1, Q.g4) : ps_w1d(xb, x, W, Q.g1, Q.g4);
    s_wx[e] = w;
  }
  __syncthreads();  // s_bar is initialised
  pdl_wait();  // (see pairwise_dual_kernel: the prologue above runs under the previous kernel's tail)
  pdl_launch_dependents();
  K.scale2 = Q.kappa4 * (Q.p.grad_out ? __ldg(Q.p.grad_out + (Q.p.per_image ? K.b : 0)) : 1.f);

  // ---- each warp: its own rows (the centre rows of its two segments; warp 3 also the 2 look-ahead rows) ----
  {
    const int r0 = min(2 * warp * S, rows), r1 = warp == PS_WARPS - 1 ? rows : min(2 * (warp + 1) * S, rows);
    const int re = min(r0 + 2, r1);  // the two rows the warp above looks ahead into
    if (Q.use_tma) {
      // try_wait suspends the warp in hardware (up to the hint) instead of spinning: waiting warps must not take
      // issue slots from the CTAs that are marching on the same SM
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
              ps_smem_u32(&s_bar)),
          "r"(20000u)
          : "memory");
    } else {
      ps_rows_load_slow<C>(Q, K, s_img, s_p, r0, r1, lane);
      __syncwarp();
    }
    PS_TR(1);
    if (Q.use_tma && tid == 0) {
      // This CTA's tile has landed: pull the tile of the CTA that will take over an SM slot one wave from now into
      // L2, so that its load is an L2 hit instead of a second trip to HBM behind everybody else's.
      const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      const unsigned nxt = cta + (unsigned)WSDL_NUM_SMS * PsCfg<C, SOFTMAX>::CTAS;
      if (nxt < gridDim.x * gridDim.y * gridDim.z) {
        const unsigned rb = nxt % gridDim.x, rest = nxt / gridDim.x;
        const unsigned tx = rest % gridDim.y, nb_img = rest / gridDim.y;
        const int ys2 = (int)rb * (PS_SEGS * S - 2);
#pragma unroll
        for (int c = 0; c < 3 + C; ++c) {
          const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
          const int pl = c < 3 ? (int)nb_img * 3 + c : (int)nb_img * C + (c - 3);
          asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm),
                       "r"((int)tx * PS_TW - 4), "r"(ys2 - 2), "r"(pl)
                       : "memory");
        }
      }
    }
#ifndef WSDL_X_NOTRANSFORM
    ps_rows_transform<C, CS, SOFTMAX>(Q, K, s_img, s_p, r0, re, lane);
#endif
    if (warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp) : "memory");  // pairs with the bar.sync of warp - 1
    PS_TR(2);
#ifndef WSDL_X_NOTRANSFORM
    ps_rows_transform<C, CS, SOFTMAX>(Q, K, s_img, s_p, re, r1, lane);
#endif
    __syncwarp();
    PS_TR(3);
    if (warp < PS_WARPS - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
  }
  PS_TR(4);

  // ---- march ----
  const int t0 = seg * S, t1 = min(t0 + S, K.nc);
  float oy[4][CS], oz[4][CS];  // what this segment adds to the first two rows of the next one
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < CS; ++c) oy[q][c] = 0.f, oz[q][c] = 0.f;
#ifdef WSDL_X_NOMARCH
  if (false) {
#else
  if (warp * 2 * S < K.nc) {  // warp-uniform: at least one of its two segments has rows
#endif
#if WSDL_PS_PACKED
    float2 A[4][CS], Bq[4][CS], Cq[4][CS];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int c = 0; c < CS; ++c) A[w][c] = Bq[w][c] = Cq[w][c] = make_float2(0.f, 0.f);
#else
    float A[8][CS], Bq[8][CS], Cq[8][CS];
    ps_zero<CS>(A), ps_zero<CS>(Bq), ps_zero<CS>(Cq);
#endif
    // one copy of the step in the instruction stream (the body is ~11 KB); the accumulator rows rotate by moves
#pragma unroll 1
    for (int s = 0; s < S; ++s) {
      const int t = t0 + s;
      const bool act = t < t1;
      float pc[4][CS], own[4][CS];
      if (act) {
        // exponent offsets of this centre row: spatial term + multiplicity of the row pairs at the top / bottom border
        const int y = K.ys - 2 + t;
        const float l0 = (y == 1 || y == H - 2) ? Q.l1g : 0.f;
        const float l1 = (y == 0 || y == H - 2) ? Q.l32 : 0.f;
        const float l2 = (y == 0 || y == H - 3) ? Q.l32 : 0.f;
        PsKs ks;
        ks.a1 = ksu + l0, ks.a4 = 4.f * ksu + l0;
        ks.b0 = ksu + l1, ks.b1 = 2.f * ksu + l1, ks.b4 = 5.f * ksu + l1;
        ks.c0 = 4.f * ksu + l2, ks.c1 = 5.f * ksu + l2, ks.c4 = 8.f * ksu + l2;
#if WSDL_PS_PACKED
        ps_stepp<CS>(A, Bq, Cq, pc, s_img, s_p, t * PS_PITCH + 4 * strip, ks);
      }
      ps_exchangep<CS>(A, own, strip);
#else
        ps_step<CS>(A, Bq, Cq, pc, s_img, s_p, t * PS_PITCH + 4 * strip, ks);
      }
      ps_exchange<CS>(A, own, strip);
#endif
      if (act) {
        if (s >= 2) {
          ps_emit<C, CS, SOFTMAX>(Q, K, t, strip, okmask, own, pc, s_gband, lsum);
        } else if (seg > 0) {
#pragma unroll
          for (int c = 0; c < CS; ++c)
            *reinterpret_cast<float4*>(s_head + (((seg - 1) * 2 + s) * CS + c) * 64 + 4 * strip) =
                make_float4(own[0][c], own[1][c], own[2][c], own[3][c]);
        }
      }
#if WSDL_PS_PACKED
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int c = 0; c < CS; ++c) A[w][c] = Bq[w][c], Bq[w][c] = Cq[w][c], Cq[w][c] = make_float2(0.f, 0.f);
#else
#pragma unroll
      for (int w = 0; w < 8; ++w)
#pragma unroll
        for (int c = 0; c < CS; ++c) A[w][c] = Bq[w][c], Bq[w][c] = Cq[w][c], Cq[w][c] = 0.f;
#endif
#ifdef WSDL_PS_TRACE
      if (s < 4) PS_TR(5 + s);
#endif
    }
    // rows t0+S, t0+S+1 belong to the next segment: what this one contributed to them is added after the barrier
#if WSDL_PS_PACKED
    ps_exchangep<CS>(A, oy, strip);
    ps_exchangep<CS>(Bq, oz, strip);
#else
    ps_exchange<CS>(A, oy, strip);
    ps_exchange<CS>(Bq, oz, strip);
#endif
  }
  PS_TR(9);
  __syncthreads();  // every head row holds its own segment's part
  if (seg < PS_SEGS - 1) {  // exactly one thread adds to each element: the strip of the segment above
#pragma unroll
    for (int c = 0; c < CS; ++c) {
      if (t0 + S < K.nc) {
        float4* h = reinterpret_cast<float4*>(s_head + ((seg * 2 + 0) * CS + c) * 64 + 4 * strip);
        float4 v = *h;
        v.x += oy[0][c], v.y += oy[1][c], v.z += oy[2][c], v.w += oy[3][c];
        *h = v;
      }
      if (t0 + S + 1 < K.nc) {
        float4* h = reinterpret_cast<float4*>(s_head + ((seg * 2 + 1) * CS + c) * 64 + 4 * strip);
        float4 v = *h;
        v.x += oz[0][c], v.y += oz[1][c], v.z += oz[2][c], v.w += oz[3][c];
        *h = v;
      }
    }
  }
  __syncthreads();
  PS_TR(10);

  // ---- the first two rows of segments 1..7, now complete ----
#ifdef WSDL_X_NOTAIL
  if (false) {
#else
  if (seg > 0) {
#endif
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int t = t0 + i;
      if (t < t1) {
        float G[4][CS], pc[4][CS];
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          const float4 h = *reinterpret_cast<const float4*>(s_head + (((seg - 1) * 2 + i) * CS + c) * 64 + 4 * strip);
          G[0][c] = h.x, G[1][c] = h.y, G[2][c] = h.z, G[3][c] = h.w;
          const float2 p01 = *reinterpret_cast<const float2*>(s_p + c * PS_PLANE + t * PS_PITCH + 4 * strip + 2);
          const float2 p23 = *reinterpret_cast<const float2*>(s_p + c * PS_PLANE + t * PS_PITCH + 4 * strip + 4);
          pc[0][c] = p01.x, pc[1][c] = p01.y, pc[2][c] = p23.x, pc[3][c] = p23.y;
        }
        ps_emit<C, CS, SOFTMAX>(Q, K, t, strip, okmask, G, pc, s_gband, lsum);
      }
    }
  }

  PS_TR(11);
  // ---- band columns (and corners): G + weight correction -> final gradient of these pixels (stored again) ----
#ifdef WSDL_X_NOTAIL
  if (false) {
#else
  if (K.xband) {
#endif
    __syncthreads();  // s_gband is complete; the first stores of these pixels are ordered before the ones below
    const int nlo = max(0, min(3, xe) - K.x0), hi0 = max(W - 3, K.x0), nhi = max(0, xe - hi0);
    const int ncb = nlo + nhi;
    const size_t plane = (size_t)H * W;
    for (int i = tid; i < ncb * K.n; i += PS_THREADS) {  // column fastest
      const int ty = i / ncb, k = i - ty * ncb;
      const int x = k < nlo ? K.x0 + k : hi0 + (k - nlo), y = K.ys + ty;
      const int slot = ps_band_slot(x, W);
      float acc[CS], pz[CS], o[C];
      if (H >= 10) {
        // (true pair weight - weight the march applied) / 2 * k (p(a) - p(b)) over the five rows of the pixel's two
        // special partner columns; column and row weights from the tables
        const int so = (ty + 2) * PS_PITCH + (x - (K.x0 - 4));
        const float i0 = s_img[so], i1 = s_img[PS_PLANE + so], i2 = s_img[2 * PS_PLANE + so];
#pragma unroll
        for (int c = 0; c < CS; ++c) pz[c] = s_p[c * PS_PLANE + so], acc[c] = 0.f;
        const float* wy = s_wx + 48 + (y < 5 ? y : (y > H - 6 ? y - (H - 10) : 10)) * 20;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 fx = *reinterpret_cast<const float4*>(s_wx + (slot * 2 + u) * 4);
          const int sj = so + __float_as_int(fx.x) - 2;
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const float4 w = *reinterpret_cast<const float4*>(wy + r * 4);
            const int sn = sj + (r - 2) * PS_PITCH;  // rows outside the image: weights 0, the staged values are finite
            const float d0 = i0 - s_img[sn], d1 = i1 - s_img[PS_PLANE + sn], d2 = i2 - s_img[2 * PS_PLANE + sn];
            const float kk = fmaf(w.x, fx.y, fmaf(w.y, fx.z, -fx.w * w.z)) * ex2_approx(fmaf(-d2, d2, fmaf(-d1, d1, -d0 * d0)));
#pragma unroll
            for (int c = 0; c < CS; ++c) acc[c] = fmaf(kk, pz[c] - s_p[c * PS_PLANE + sn], acc[c]);
          }
        }
      } else {
        ps_xfix_item<CS>(H, Q.g1, Q.g4, 1.f, s_img, s_p, s_wx + slot * 10, K.ys, K.x0, y, x, acc, pz);
      }
      lsum += ps_pixel_grad<C, CS, SOFTMAX>(K.scale2, pz, acc, o);  // the loss is linear in G: add the correction's share
      if (Q.p.grad_values) {
#pragma unroll
        for (int c = 0; c < CS; ++c) acc[c] += s_gband[(slot * CS + c) * PS_CAP + ty];
        ps_pixel_grad<C, CS, SOFTMAX>(K.scale2, pz, acc, o);
        float* go = Q.p.grad_values + (size_t)K.b * C * plane + (size_t)y * W + x;
#pragma unroll
        for (int c = 0; c < C; ++c) go[c * plane] = o[c];
      }
    }
  }

  PS_TR(12);
  // ---- one loss partial per CTA (fire and forget, one 32-bit store into the CTA's own slot); the last CTA of the grid
  // polls the slots until none holds the "empty" pattern (all ones: a NaN no arithmetic produces), adds them per image in
  // a fixed order, in double, and puts the pattern back.  No ticket, no fence: a CTA retires without waiting for the L2
  // to acknowledge its partial.  Every other CTA was dispatched before the finisher and none waits for it. ----
  const int kpi = Q.nb * Q.n_x;  // CTAs (= partials) per image
  const bool finisher = blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1;
  {
    const float w = warp_sum(lsum);
    if (lane == 0) s_red[warp] = w;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < PS_THREADS / 32; ++i) t += s_red[i];
      unsigned v = __float_as_uint(2.f * t);
      if (v == 0xffffffffu) v = 0x7fffffffu;  // cannot happen with finite inputs; keep a NaN a NaN
      __stcg(Q.slots + (size_t)K.b * kpi + blockIdx.y * Q.nb + blockIdx.x, v);
    }
  }
  PS_TR(13);
#ifdef WSDL_PS_TRACE
  {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (lane == 0 && cta < 4096) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      ps_trace_buf[(cta * 4 + warp) * 16 + 15] = smid;
    }
  }
#endif
  if (!finisher) return;
  double wtot = 0.0;
  for (int b = warp; b < Q.p.B; b += PS_THREADS / 32) {
    double acc = 0.0;
    for (int i = lane; i < kpi; i += 32) {
      unsigned* slot = Q.slots + (size_t)b * kpi + i;
      unsigned v, spins = 0;
      while (true) {
        asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(slot) : "memory");
        if (v != 0xffffffffu) break;
        __nanosleep(100);
        if (++spins > (1u << 24)) __trap();  // a CTA that never reports: fail instead of hanging the device
      }
      __stcg(slot, 0xffffffffu);
      acc += (double)__uint_as_float(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (Q.p.per_image) {
      if (lane == 0) Q.p.loss_out[b] = (float)(acc * Q.p.kappa);
    } else {
      wtot += acc;
    }
  }
  if (!Q.p.per_image) {
    if (lane == 0) s_dred[warp] = wtot;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
#pragma unroll
      for (int i = 0; i < PS_THREADS / 32; ++i) t += s_dred[i];
      Q.p.loss_out[0] = (float)(t * Q.p.kappa);
    }
  }
}

// 4 finished pixels of centre row t: G[.][0] cut, G[.][1] boundary, one probability p0 each
__device__ __forceinline__ void ps_emit_dual(const PsParams& Q, const PsBlk& K, float scale_b, int t, int strip, int okmask,
                                             const float (&G)[4][2], const float (&pc)[4], float* gband_row, float& lsum_c,
                                             float& lsum_b) {
  const int H = Q.p.H, W = Q.p.W;
  const int y = K.ys - 2 + t;
  const int xs = K.x0 - 2 + 4 * strip;
  if (okmask >> 4) {  // this lane owns band-column pixels (bits 4.. : their slots + 1): the band pass finishes them from G
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int slot1 = (okmask >> (4 + 4 * j)) & 15;
      if (slot1) *reinterpret_cast<float2*>(gband_row + (slot1 - 1) * 2) = make_float2(G[j][0], G[j][1]);
    }
  }
  float out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float p0 = pc[j], p1 = 1.f - p0;
    out[j] = 2.f * p0 * p1 * fmaf(K.scale2, G[j][0], scale_b * G[j][1]);  // dL/dlogit0 = -dL/dlogit1
    if ((okmask >> j) & 1) {
      lsum_c = fmaf(p0 - p1, G[j][0], lsum_c);
      lsum_b = fmaf(p0 - p1, G[j][1], lsum_b);
    }
  }
  if (Q.p.grad_values) {
    const size_t plane = (size_t)H * W;
    float* go = Q.p.grad_values + (size_t)K.b * 2 * plane + (size_t)y * W + xs;
    if (Q.vec2_ok) {
      if (okmask & 1) {
        *reinterpret_cast<float2*>(go) = make_float2(out[0], out[1]);
        *reinterpret_cast<float2*>(go + plane) = make_float2(-out[0], -out[1]);
      }
      if (okmask & 4) {
        *reinterpret_cast<float2*>(go + 2) = make_float2(out[2], out[3]);
        *reinterpret_cast<float2*>(go + plane + 2) = make_float2(-out[2], -out[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if ((okmask >> j) & 1) go[j] = out[j], go[plane + j] = -out[j];
    }
  }
}

// Geometry of the dual kernel: segments of up to 6 rows (48 centre rows, 46 owned).  A block's fixed costs -- tile wait,
// prologue, segment heads, loss ticket -- are paid once per CTA, and 224 rows split into 5 blocks instead of 6.  The tile
// grows to 68.5 KB; three CTAs still share an SM because the segment heads and the band-column G no longer have a region
// of their own: they live in shared memory that is dead by the time they are written (below).
constexpr int DU_SMAX = 6;
constexpr int DU_ROWS = PS_SEGS * DU_SMAX + 2;
constexpr int DU_CAP = PS_SEGS * DU_SMAX - 2;
constexpr int DU_PLANE = (DU_ROWS * PS_PITCH + 31) / 32 * 32;
constexpr size_t PS_DUAL_SMEM_FLOATS = (size_t)5 * DU_PLANE + 2 * 6 * 10 + 6 * 2 * 8 + 11 * 5 * 8;

__global__ void __launch_bounds__(PS_THREADS, WSDL_PS_CTAS)
    pairwise_dual_kernel(const __grid_constant__ PsParams Q, const __grid_constant__ PsDual D,
                         const __grid_constant__ CUtensorMap tm_img, const __grid_constant__ CUtensorMap tm_val) {
  constexpr int C = 2;
  extern __shared__ __align__(128) float ps_smem[];
  float* s_img = ps_smem;                                // [3][DU_ROWS][PS_PITCH]: raw, then scaled for sigma_cut
  float* s_p = s_img + 3 * DU_PLANE;                     // [2][DU_ROWS][PS_PITCH]: logits, then p0 in plane 0
  float* s_wx = s_p + C * DU_PLANE;                      // [2][6][2][5]: column weights, cut then boundary
  float* s_fx = s_wx + 2 * 6 * 10;                       // [6][2][8]: per band slot, its two special partner columns
  float* s_wy = s_fx + 6 * 2 * 8;                        // [11][5][8]: row weights of the 10 row-band rows + the interior
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ float s_red[2][PS_THREADS / 32];
  __shared__ double s_dred[PS_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = warp * 2 + (lane >> 4), strip = lane & 15;
  const int H = Q.p.H, W = Q.p.W;
  const int S = Q.S;

  PsBlk K;
  K.b = blockIdx.z;
  K.x0 = blockIdx.y * PS_TW;
  K.ys = blockIdx.x * (PS_SEGS * S - 2);
  K.n = min(PS_SEGS * S - 2, H - K.ys);
  K.nc = K.n + 2;
  const int xe = min(K.x0 + PS_TW, W);
  K.xband = (K.x0 <= 2) || (xe - 1 >= W - 3);
  const int rows = K.nc + 2;
  // Per warp, 512 + 24 S floats: G (cut, boundary) of the first two rows of its two segments [h][i][c][64], then G of
  // the band-column pixels of its 2 S centre rows [row][6 slots][cut, boundary].  With S >= 5 they fit into the warp's
  // OWN rows of the second logit plane -- dead once the warp has converted them (p0 lives in plane 0), which it has
  // before it marches, so there is no block-wide barrier between conversion and march; shorter blocks leave the rows
  // past 8 S + 2 of every plane unused, and warp w takes those of plane w.
  const auto wreg = [&](int w) -> float* {
    return S >= 5 ? s_p + DU_PLANE + 2 * w * S * PS_PITCH : s_img + w * DU_PLANE + (PS_SEGS * S + 2) * PS_PITCH;
  };
  float* const my_head = wreg(warp) + (lane >> 4) * 256;  // + (i * 2 + c) * 64 + 4 * strip
  float* const my_gband = wreg(warp) + 512 - 2 * warp * S * 12;  // + t * 12 + slot * 2 for a centre row t of this warp
  PS_TR(0);
#ifdef WSDL_PS_TRACE
  if (lane == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned c_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (c_ < 4096) ps_trace_buf[(c_ * 4 + warp) * 16 + 15] = smid;
  }
#endif

  if (Q.use_tma && tid == 0) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const unsigned bar = ps_smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_wait();  // programmatic dependent launch: nothing before this line touches global memory
    if (cta >= WSDL_NUM_SMS && cta < (unsigned)WSDL_NUM_SMS * WSDL_PS_CTAS && Q.stagger_ns > 0)
      __nanosleep((cta / WSDL_NUM_SMS) * Q.stagger_ns);
    const unsigned bytes = (unsigned)(5 * (PS_SEGS * S + 2) * PS_PITCH * sizeof(float));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#pragma unroll
    for (int c = 0; c < 5; ++c) {
      const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
      const int pl = c < 3 ? K.b * 3 + c : K.b * C + (c - 3);
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
              ps_smem_u32(s_img + c * DU_PLANE)),
          "l"(tm), "r"(K.x0 - 4), "r"(K.ys - 2), "r"(pl), "r"(bar)
          : "memory");
    }
  }
  int okmask = 0;  // bits 0..3: pixel j of the strip is owned; bits 4 + 4 j ..: its band slot + 1 (0: not a band column)
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = 4 * strip + j;
    const bool ok = col >= 2 && col < 2 + PS_TW && K.x0 - 2 + col < W;
    if (ok) okmask |= 1 << j;
    const int slot = ok ? ps_band_slot(K.x0 - 2 + col, W) : -1;
    if (slot >= 0) okmask |= (slot + 1) << (4 + 4 * j);
  }
  float lsum_c = 0.f, lsum_b = 0.f;
  if (K.xband) {
    if (H >= 10) {  // the band pass's weight tables: built on the host once per launch (ps_band_tables), 2.1 KB
      for (int i = tid; i < 6 * 2 * 8 + 11 * 5 * 8; i += PS_THREADS) s_fx[i] = i < 96 ? D.fx[i] : D.wy[i - 96];
    } else if (tid < 120) {  // tiny images take the generic band pass: its column weights for gamma = 1 and gamma_b
      const int which = tid / 60, e = tid - which * 60, slot = e / 10, rem = e - slot * 10, j = rem % 5;
      const float g1 = which ? D.g1b : 1.f, g4 = which ? D.g4b : 1.f;
      const int x = slot < 3 ? slot : W - 6 + slot, xb = x + j - 2;
      float w = 0.f;
      if (xb >= 0 && xb < W) w = rem < 5 ? ps_w1d(x, xb, W, g1, g4) : ps_w1d(xb, x, W, g1, g4);
      s_wx[tid] = w;
    }
  }
  __syncthreads();  // s_bar is initialised
  // Everything above is index arithmetic and shared-memory tables: launched with programmatic stream serialization, a
  // CTA runs it while the previous kernel of the stream is still draining, and waits here (tid 0: before its tile load)
  // for that kernel's memory to be visible.  Dependents of THIS launch may be scheduled as soon as every CTA got here.
  pdl_wait();
  pdl_launch_dependents();
  K.scale2 = Q.kappa4 * (Q.p.grad_out ? __ldg(Q.p.grad_out) : 1.f);
  const float scale_b = D.kappa_bnd4 * (D.grad_out_bnd ? __ldg(D.grad_out_bnd + K.b) : 1.f);

  {
    const int r0 = min(2 * warp * S, rows), r1 = warp == PS_WARPS - 1 ? rows : min(2 * (warp + 1) * S, rows);
    const int re = min(r0 + 2, r1);
    if (Q.use_tma) {
      asm volatile(
          "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0, %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(
              ps_smem_u32(&s_bar)),
          "r"(20000u)
          : "memory");
    } else {
      ps_rows_load_slow<C, DU_PLANE>(Q, K, s_img, s_p, r0, r1, lane);
      __syncwarp();
    }
    PS_TR(1);
    if (Q.use_tma && tid == 0) {  // as in pairwise_sym_kernel: pull the tile that takes over an SM slot a wave later into L2
      const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      const unsigned nxt = cta + (unsigned)WSDL_NUM_SMS * WSDL_PS_CTAS;
      if (nxt < gridDim.x * gridDim.y * gridDim.z) {
        const unsigned rb = nxt % gridDim.x, rest = nxt / gridDim.x;
        const unsigned tx = rest % gridDim.y, nb_img = rest / gridDim.y;
        const int ys2 = (int)rb * (PS_SEGS * S - 2);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const CUtensorMap* tm = c < 3 ? &tm_img : &tm_val;
          const int pl = c < 3 ? (int)nb_img * 3 + c : (int)nb_img * C + (c - 3);
          asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tm),
                       "r"((int)tx * PS_TW - 4), "r"(ys2 - 2), "r"(pl)
                       : "memory");
        }
      }
    }
#if !(WSDL_X_SKIP & 1)
    ps_rows_transform<C, 1, true, DU_PLANE>(Q, K, s_img, s_p, r0, re, lane);
#endif
    if (warp > 0) asm volatile("bar.arrive %0, 64;" ::"r"(warp) : "memory");
    PS_TR(2);
#if !(WSDL_X_SKIP & 1)
    ps_rows_transform<C, 1, true, DU_PLANE>(Q, K, s_img, s_p, re, r1, lane);
#endif
    __syncwarp();
    PS_TR(3);
    if (warp < PS_WARPS - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
  }
  PS_TR(4);

  // ---- march ----
  const int t0 = seg * S, t1 = min(t0 + S, K.nc);
  float oy[4][2], oz[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) oy[q][0] = oy[q][1] = oz[q][0] = oz[q][1] = 0.f;
  if (!(WSDL_X_SKIP & 4) && warp * 2 * S < K.nc) {
    float2 A[4][2], Bq[4][2], Cq[4][2];
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int c = 0; c < 2; ++c) A[w][c] = Bq[w][c] = Cq[w][c] = make_float2(0.f, 0.f);
    const float ksu = D.ksu_b, ratio = D.ratio;
    // one copy of the step in the instruction stream: unrolling by 2 / 3 (which would let the accumulator rows rotate
    // by renaming instead of moves) costs 29 -> 36 / 41 us per step -- the ~13 KB body already strains the
    // instruction caches
#pragma unroll 1
    for (int s = 0; s < S; ++s) {
      const int t = t0 + s;
      const bool act = t < t1;
      float pc[4], own[4][2];
      if (act) {
        const int y = K.ys - 2 + t;
        const bool r1 = (y == 1 || y == H - 2);
        const float l0c = r1 ? 1.f : 0.f, l0b = r1 ? D.l1g_b : 0.f;  // log2(1 + gamma^4): gamma = 1 for the cut loss
        const float l1 = (y == 0 || y == H - 2) ? Q.l32 : 0.f;
        const float l2 = (y == 0 || y == H - 3) ? Q.l32 : 0.f;
        PsKsDual ks;
        ks.ca = l0c, ks.cb = l1, ks.cc = l2;
        const float ra = ratio * l0c, rb = ratio * l1, rc = ratio * l2;
        ks.a1 = ksu + l0b - ra, ks.a4 = 4.f * ksu + l0b - ra;
        ks.b0 = ksu + l1 - rb, ks.b1 = 2.f * ksu + l1 - rb, ks.b4 = 5.f * ksu + l1 - rb;
        ks.c0 = 4.f * ksu + l2 - rc, ks.c1 = 5.f * ksu + l2 - rc, ks.c4 = 8.f * ksu + l2 - rc;
        ps_step_dual2<DU_PLANE>(A, Bq, Cq, pc, s_img, s_p, t * PS_PITCH + 4 * strip, ks, ratio);
      }
      ps_exchangep<2>(A, own, strip);
      if (act) {
        if (s >= 2) {
          ps_emit_dual(Q, K, scale_b, t, strip, okmask, own, pc, my_gband + t * 12, lsum_c, lsum_b);
        } else if (seg > 0) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
            *reinterpret_cast<float4*>(my_head + (s * 2 + c) * 64 + 4 * strip) =
                make_float4(own[0][c], own[1][c], own[2][c], own[3][c]);
        }
      }
      if (s < 4) PS_TR(5 + s);
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int c = 0; c < 2; ++c) A[w][c] = Bq[w][c], Bq[w][c] = Cq[w][c], Cq[w][c] = make_float2(0.f, 0.f);
    }
    ps_exchangep<2>(A, oy, strip);
    ps_exchangep<2>(Bq, oz, strip);
  }
  PS_TR(9);
  __syncthreads();  // every head row holds its own segment's part
  PS_TR(10);
  if (seg < PS_SEGS - 1) {
    float* const next_head = wreg((seg + 1) >> 1) + ((seg + 1) & 1) * 256;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      if (t0 + S < K.nc) {
        float4* h = reinterpret_cast<float4*>(next_head + c * 64 + 4 * strip);
        float4 v = *h;
        v.x += oy[0][c], v.y += oy[1][c], v.z += oy[2][c], v.w += oy[3][c];
        *h = v;
      }
      if (t0 + S + 1 < K.nc) {
        float4* h = reinterpret_cast<float4*>(next_head + (2 + c) * 64 + 4 * strip);
        float4 v = *h;
        v.x += oz[0][c], v.y += oz[1][c], v.z += oz[2][c], v.w += oz[3][c];
        *h = v;
      }
    }
  }
  __syncthreads();

  // ---- the first two rows of segments 1..7, now complete ----
  if (!(WSDL_X_SKIP & 2) && seg > 0) {
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int t = t0 + i;
      if (t < t1) {
        float G[4][2], pc[4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4 h = *reinterpret_cast<const float4*>(my_head + (i * 2 + c) * 64 + 4 * strip);
          G[0][c] = h.x, G[1][c] = h.y, G[2][c] = h.z, G[3][c] = h.w;
        }
        const float2 p01 = *reinterpret_cast<const float2*>(s_p + t * PS_PITCH + 4 * strip + 2);
        const float2 p23 = *reinterpret_cast<const float2*>(s_p + t * PS_PITCH + 4 * strip + 4);
        pc[0] = p01.x, pc[1] = p01.y, pc[2] = p23.x, pc[3] = p23.y;
        ps_emit_dual(Q, K, scale_b, t, strip, okmask, G, pc, my_gband + t * 12, lsum_c, lsum_b);
      }
    }
  }

  PS_TR(11);
  // ---- band columns (and corners) ----
  if (!(WSDL_X_SKIP & 8) && K.xband) {
    __syncthreads();
    const int nlo = max(0, min(3, xe) - K.x0), hi0 = max(W - 3, K.x0), nhi = max(0, xe - hi0);
    const int ncb = nlo + nhi;
    const size_t plane = (size_t)H * W;
    for (int i = tid; i < ncb * K.n; i += PS_THREADS) {
      const int ty = i / ncb, k = i - ty * ncb;
      const int x = k < nlo ? K.x0 + k : hi0 + (k - nlo), y = K.ys + ty;
      const int slot = ps_band_slot(x, W);
      float ac, ab, p0;
      if (H >= 10) {
        // (true pair weight - weight the march applied) / 2 * k (p(a) - p(b)) over the five rows of the two special
        // partner columns, both losses from one set of loads and one squared distance; weights from the two tables
        const int so = (ty + 2) * PS_PITCH + (x - (K.x0 - 4));
        const float i0 = s_img[so], i1 = s_img[DU_PLANE + so], i2 = s_img[2 * DU_PLANE + so];
        p0 = s_p[so];
        const float* wy = s_wy + (y < 5 ? y : (y > H - 6 ? y - (H - 10) : 10)) * 40;
        ac = 0.f, ab = 0.f;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 fc = *reinterpret_cast<const float4*>(s_fx + (slot * 2 + u) * 8);
          const float4 fb = *reinterpret_cast<const float4*>(s_fx + (slot * 2 + u) * 8 + 4);
          const int sj = so + __float_as_int(fc.x) - 2;
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const float4 wa = *reinterpret_cast<const float4*>(wy + r * 8);
            const float2 wb2 = *reinterpret_cast<const float2*>(wy + r * 8 + 4);
            const int sn = sj + (r - 2) * PS_PITCH;  // rows outside the image: weights 0, the staged values are finite
            const float d0 = i0 - s_img[sn], d1 = i1 - s_img[DU_PLANE + sn], d2 = i2 - s_img[2 * DU_PLANE + sn];
            const float e = fmaf(-d2, d2, fmaf(-d1, d1, -d0 * d0));
            const float dp = p0 - s_p[sn];
            const float dc = fmaf(wa.x, fc.y, fmaf(wa.y, fc.z, -fc.w * wa.z));
            const float db = fmaf(wa.w, fb.x, fmaf(wb2.x, fb.y, -fb.z * wb2.y));
            ac = fmaf(dc * ex2_approx(e), dp, ac);
            ab = fmaf(db * ex2_approx(D.ratio * e), dp, ab);
          }
        }
      } else {
        ps_xfix_dual<DU_PLANE>(H, D.g1b, D.g4b, D.ratio, s_img, s_p, s_wx + slot * 10, s_wx + 60 + slot * 10, K.ys, K.x0, y, x, ac, ab, p0);
      }
      const float p1 = 1.f - p0;
      lsum_c = fmaf(p0 - p1, ac, lsum_c);  // the losses are linear in G: the corrections' share
      lsum_b = fmaf(p0 - p1, ab, lsum_b);
      if (Q.p.grad_values) {
        const int tw = min((ty + 2) / (2 * S), PS_WARPS - 1);  // the warp that marched centre row ty + 2
        const float2 g2 = *reinterpret_cast<const float2*>(wreg(tw) + 512 + (ty + 2 - 2 * tw * S) * 12 + slot * 2);
        const float gc = ac + g2.x, gb = ab + g2.y;
        const float o = 2.f * p0 * p1 * fmaf(K.scale2, gc, scale_b * gb);
        float* go = Q.p.grad_values + (size_t)K.b * 2 * plane + (size_t)y * W + x;
        go[0] = o, go[plane] = -o;
      }
    }
  }

  PS_TR(12);
  // ---- loss partials (cut, boundary) per CTA; the last CTA of the grid finishes both ----
  // No ticket and no fence: a CTA publishes its two partials as ONE naturally aligned 64-bit store into its own slot, and
  // the finisher polls the slots until none holds the "empty" pattern (all ones: a NaN no arithmetic produces), sums
  // them in a fixed order and puts the pattern back.  A release reduction on a ticket would make every CTA wait for the
  // L2 to acknowledge its partials before it may retire -- 0.4 us of a fused launch (profiles/r02_g_phase_skip.txt).
  const int kpi = Q.nb * Q.n_x;
  const bool finisher = blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1 && blockIdx.z == gridDim.z - 1;
  {
    const float wc = warp_sum(lsum_c), wb = warp_sum(lsum_b);
    if (lane == 0) s_red[0][warp] = wc, s_red[1][warp] = wb;
    __syncthreads();
    if (tid == 0) {
      float tc = 0.f, tb = 0.f;
#pragma unroll
      for (int i = 0; i < PS_THREADS / 32; ++i) tc += s_red[0][i], tb += s_red[1][i];
      unsigned long long v = (unsigned long long)__float_as_uint(2.f * tc) | ((unsigned long long)__float_as_uint(2.f * tb) << 32);
      if (v == PS_SLOT_EMPTY) v = 0x7fffffff7fffffffull;  // cannot happen with finite inputs; keep a NaN a NaN
      __stcg(D.slots + (size_t)K.b * kpi + blockIdx.y * Q.nb + blockIdx.x, v);
    }
  }
  PS_TR(13);
  if (!finisher) return;
  double wtot = 0.0;
  for (int b = warp; b < Q.p.B; b += PS_THREADS / 32) {  // image by image: cut summed over the batch, boundary per image
    double ac = 0.0, ab = 0.0;
    for (int i = lane; i < kpi; i += 32) {
      unsigned long long* slot = D.slots + (size_t)b * kpi + i;
      unsigned long long v;
      unsigned spins = 0;
      while ((v = ps_ld_slot(slot)) == PS_SLOT_EMPTY) {
        __nanosleep(100);
        if (++spins > (1u << 24)) __trap();  // a CTA that never reports: fail instead of hanging the device
      }
      __stcg(slot, PS_SLOT_EMPTY);
      ac += (double)__uint_as_float((unsigned)v);
      ab += (double)__uint_as_float((unsigned)(v >> 32));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ac += __shfl_xor_sync(0xffffffffu, ac, o), ab += __shfl_xor_sync(0xffffffffu, ab, o);
    if (lane == 0) D.loss_bnd[b] = (float)(ab * D.kappa_bnd);
    wtot += ac;
  }
  if (lane == 0) s_dred[warp] = wtot;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < PS_THREADS / 32; ++i) t += s_dred[i];
    Q.p.loss_out[0] = (float)(t * Q.p.kappa);
  }
}

// Row blocks per column tile: at most PS_CAP rows each; among the next few candidates the count with the lowest
// modelled time = (march steps per block + fixed cost of a block, in steps) x blocks an SM works through (a launch
// below two blocks per SM is latency bound: more, shorter blocks keep winning there).
static int ps_row_blocks(int B, int H, int W, int cap = PS_CAP) {
  static const int forced = WSDL_TUNE_INT("WSDL_PS_NB", 0);
  const int n_x = (W + PS_TW - 1) / PS_TW;
  const int nb_min = (H + cap - 1) / cap;
  if (forced >= nb_min) return forced;
  int best = nb_min;
  double best_cost = 1e300;
  for (int nb = nb_min; nb <= nb_min + 12; ++nb) {
    const int n = (H + nb - 1) / nb;  // rows of the largest block
    if (n < 6 && nb > nb_min) break;
    const int S = (n + 2 + PS_SEGS - 1) / PS_SEGS < 2 ? 2 : (n + 2 + PS_SEGS - 1) / PS_SEGS;
    const double per_sm = (double)B * n_x * nb / WSDL_NUM_SMS;
    const double rounds = per_sm < 2.0 ? 2.0 : per_sm;  // blocks of concurrent launches fill the ragged last round
    const double cost = (S + 1.5) * rounds;
    if (cost < best_cost - 1e-9) best_cost = cost, best = nb;
  }
  return best;
}

size_t ps_workspace_floats(int B, int H, int W) {  // one loss partial per block, whichever kernel cuts the rows
  const int n_x = (W + PS_TW - 1) / PS_TW;
  const int nb = ps_row_blocks(B, H, W), nb_dual = ps_row_blocks(B, H, W, DU_CAP);
  return (size_t)B * n_x * (nb > nb_dual ? nb : nb_dual);
}

typedef CUresult (*PsEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PsEncodeFn ps_encoder() {  // cuTensorMapEncodeTiled through the runtime: no link-time dependency on libcuda
  static const PsEncodeFn fn = []() -> PsEncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (PsEncodeFn)p;
  }();
  return fn;
}

// (W, H, planes) f32 tensor, box = 68 columns x `rows` rows x 1 plane, zero fill outside
static bool ps_encode(CUtensorMap* tm, const float* base, int W, int H, long long planes, int rows) {
  const PsEncodeFn enc = ps_encoder();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
  const cuuint32_t box[3] = {(cuuint32_t)PS_PITCH, (cuuint32_t)rows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int C, bool SOFTMAX>
static int ps_launch_t(PsParams& Q, const CUtensorMap& tm_img, const CUtensorMap& tm_val, cudaStream_t s) {
  constexpr size_t smem = PsCfg<C, SOFTMAX>::smem_floats * sizeof(float);
  // function attributes are per device: remember which devices have them (idempotent; a race only repeats the call)
  static bool attr_done[64] = {};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, attr_done[0] = false;
  bool& attr_set = attr_done[dev_id];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_sym_kernel<C, SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(pairwise_sym_kernel<C, SOFTMAX>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  static const size_t pad = (size_t)WSDL_TUNE_INT("WSDL_PS_SMEM_PAD_KB", 0) * 1024;
  if (pad) cudaFuncSetAttribute(pairwise_sym_kernel<C, SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + pad));
  {
    const cudaError_t e = launch_pdl(pairwise_sym_kernel<C, SOFTMAX>, dim3(Q.nb, Q.n_x, Q.p.B), dim3(PS_THREADS), smem + pad, s, Q,
                                     tm_img, tm_val);
    if (e != cudaSuccess) return (int)e;
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}

int ps_launch(const PwParams& P, cudaStream_t s) {
  if (P.pad != 2 || P.C < 1 || P.C > 2 || P.H < 6 || P.W < 6) return 1;
  PsParams Q;
  Q.p = P;
  Q.n_x = (P.W + PS_TW - 1) / PS_TW;
  Q.nb = ps_row_blocks(P.B, P.H, P.W);
  if (Q.nb > 65535 || Q.n_x > 65535 || P.B > 65535) return 1;
  const int n_max = (P.H + Q.nb - 1) / Q.nb;
  Q.S = (n_max + 2 + PS_SEGS - 1) / PS_SEGS < 2 ? 2 : (n_max + 2 + PS_SEGS - 1) / PS_SEGS;
  Q.nb = (P.H + PS_SEGS * Q.S - 3) / (PS_SEGS * Q.S - 2);  // blocks of 8 S - 2 rows: never more than the model chose
  Q.slots = P.sym_slots;
  Q.kappa4 = (float)(4.0 * P.kappa);
  Q.vec2_ok = ((P.W & 1) == 0) && (!P.grad_values || ((uintptr_t)P.grad_values & 7) == 0);
  Q.img_scale = sqrtf(-P.kc);
  Q.g1 = expf(-P.inv_2ss), Q.g4 = expf(-4.f * P.inv_2ss);
  Q.l32 = log2f(1.5f), Q.l1g = log2f(1.f + Q.g4);
  if (P.H >= 10) ps_band_tables1(Q, P.H, P.W);
  else memset(Q.fx1, 0, sizeof(Q.fx1)), memset(Q.wy1, 0, sizeof(Q.wy1));
  {  // a sixteenth of a wave of tiles at ~6 TB/s (a longer hold-back helps a lone launch by 2-3 % and costs launches
     // that overlap on the device 4 %); only when the launch fills the machine
    static const int stagger_env = WSDL_TUNE_INT("WSDL_PS_STAGGER_NS", -1);
    const double tile_bytes = (double)(3 + P.C) * (PS_SEGS * Q.S + 2) * PS_PITCH * 4.0;
    const long long ctas = (long long)Q.nb * Q.n_x * P.B;
    Q.stagger_ns = ctas >= 2LL * WSDL_NUM_SMS ? (unsigned)(tile_bytes * WSDL_NUM_SMS / 24000.0) : 0u;
    if (stagger_env >= 0) Q.stagger_ns = (unsigned)stagger_env;
  }
  CUtensorMap tm_img, tm_val;
  memset(&tm_img, 0, sizeof(tm_img)), memset(&tm_val, 0, sizeof(tm_val));
  static const int no_tma = WSDL_TUNE_INT("WSDL_PAIRWISE_NO_TMA", 0);
  Q.use_tma = !no_tma && ((P.W & 3) == 0) && (((uintptr_t)P.values & 15) == 0) && (((uintptr_t)P.images & 15) == 0) &&
              ps_encode(&tm_img, P.images, P.W, P.H, 3LL * P.B, PS_SEGS * Q.S + 2) &&
              ps_encode(&tm_val, P.values, P.W, P.H, (long long)P.C * P.B, PS_SEGS * Q.S + 2);
  if (P.C == 2)
    return P.inner_softmax ? ps_launch_t<2, true>(Q, tm_img, tm_val, s) : ps_launch_t<2, false>(Q, tm_img, tm_val, s);
  return P.inner_softmax ? ps_launch_t<1, true>(Q, tm_img, tm_val, s) : ps_launch_t<1, false>(Q, tm_img, tm_val, s);
}

// Host side of the dual kernel.  P describes the cut loss (values = logits, C = 2, inner softmax, batch-mean loss,
// kc for sigma_cut, no spatial term); the boundary loss comes in through the extra arguments.
int ps_launch_dual(const PwParams& P, float sigma_cut, float sigma_bnd, float sigma_space, const float* grad_out_bnd,
                   float* loss_bnd, unsigned long long* slots, cudaStream_t s) {
  if (P.pad != 2 || P.C != 2 || P.H < 6 || P.W < 6 || !P.inner_softmax || P.per_image) return 1;
  PsParams Q;
  Q.p = P;
  Q.n_x = (P.W + PS_TW - 1) / PS_TW;
  Q.slots = nullptr;
  Q.kappa4 = (float)(4.0 * P.kappa);
  memset(Q.fx1, 0, sizeof(Q.fx1)), memset(Q.wy1, 0, sizeof(Q.wy1));
  Q.nb = ps_row_blocks(P.B, P.H, P.W, DU_CAP);
  if (Q.nb > 65535 || Q.n_x > 65535 || P.B > 65535) return 1;
  const int n_max = (P.H + Q.nb - 1) / Q.nb;
  Q.S = (n_max + 2 + PS_SEGS - 1) / PS_SEGS < 2 ? 2 : (n_max + 2 + PS_SEGS - 1) / PS_SEGS;
  Q.nb = (P.H + PS_SEGS * Q.S - 3) / (PS_SEGS * Q.S - 2);
  Q.vec2_ok = ((P.W & 1) == 0) && (!P.grad_values || ((uintptr_t)P.grad_values & 7) == 0);
  Q.img_scale = sqrtf(-P.kc);
  Q.g1 = 1.f, Q.g4 = 1.f;
  Q.l32 = log2f(1.5f), Q.l1g = 1.f;
  {
    static const int stagger_env = WSDL_TUNE_INT("WSDL_PS_STAGGER_NS", -1);
    const double tile_bytes = 5.0 * (PS_SEGS * Q.S + 2) * PS_PITCH * 4.0;
    const long long ctas = (long long)Q.nb * Q.n_x * P.B;
    Q.stagger_ns = ctas >= 2LL * WSDL_NUM_SMS ? (unsigned)(tile_bytes * WSDL_NUM_SMS / 24000.0) : 0u;
    if (stagger_env >= 0) Q.stagger_ns = (unsigned)stagger_env;
  }
  PsDual D;
  D.grad_out_bnd = grad_out_bnd;
  D.loss_bnd = loss_bnd;
  D.slots = slots;
  D.kappa_bnd = 1.0 / (24.0 * (double)P.H * (double)P.W);
  D.kappa_bnd4 = (float)(4.0 * D.kappa_bnd);
  D.ratio = (sigma_cut * sigma_cut) / (sigma_bnd * sigma_bnd);
  const float inv_2ss = sigma_space > 0.f ? 1.f / (2.f * sigma_space * sigma_space) : 0.f;
  D.ksu_b = -LOG2E * inv_2ss;
  D.g1b = expf(-inv_2ss), D.g4b = expf(-4.f * inv_2ss);
  D.l1g_b = log2f(1.f + D.g4b);
  if (P.H >= 10) ps_band_tables(D, P.H, P.W);
  else memset(D.fx, 0, sizeof(D.fx)), memset(D.wy, 0, sizeof(D.wy));
  CUtensorMap tm_img, tm_val;
  memset(&tm_img, 0, sizeof(tm_img)), memset(&tm_val, 0, sizeof(tm_val));
  static const int no_tma = WSDL_TUNE_INT("WSDL_PAIRWISE_NO_TMA", 0);
  Q.use_tma = !no_tma && ((P.W & 3) == 0) && (((uintptr_t)P.values & 15) == 0) && (((uintptr_t)P.images & 15) == 0) &&
              ps_encode(&tm_img, P.images, P.W, P.H, 3LL * P.B, PS_SEGS * Q.S + 2) &&
              ps_encode(&tm_val, P.values, P.W, P.H, 2LL * P.B, PS_SEGS * Q.S + 2);
  constexpr size_t smem = PS_DUAL_SMEM_FLOATS * sizeof(float);
  static bool attr_done[64] = {};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0, attr_done[0] = false;
  if (!attr_done[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(pairwise_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(pairwise_dual_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                             (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    attr_done[dev_id] = true;
  }
  {
    const cudaError_t e = launch_pdl(pairwise_dual_kernel, dim3(Q.nb, Q.n_x, P.B), dim3(PS_THREADS), smem, s, Q, D, tm_img, tm_val);
    if (e != cudaSuccess) return (int)e;
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}

}  // namespace wsdl


#ifdef WSDL_PS_TRACE
extern "C" int wsdl_ps_trace_read(unsigned long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, wsdl::ps_trace_buf, (size_t)n * 64 * sizeof(unsigned long long));
}
#endif
