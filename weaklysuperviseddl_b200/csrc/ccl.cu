// keep_largest: 8-connected component labelling + largest-area selection on the GPU, sm_100a.
//
// Replaces keep_largest (reference TraditionalModel/PsuedoMasks.py:15-21; duplicate at
// AlternatingDirectionCutLoss.py:206-213): skimage.measure.label (full connectivity, background 0),
// regionprops, max by area (first maximum = lowest label = component whose first pixel comes first
// in raster order), `labeled == label`.  Empty masks are returned unchanged.
//
// Union-find with the minimum pixel index as root (so root order == skimage label order), in two levels:
//   1. ccl_tile    one CTA labels a 32x32 tile in SHARED memory: horizontal runs are resolved with one ballot per
//                  row (a pixel starts at the first pixel of its run), runs are joined to the row above with shared-
//                  memory atomicMin unions (only where the left neighbour has not already made the same join), the
//                  tile is flattened, pixels are counted per local root, and every pixel writes the GLOBAL index of
//                  its local root; local roots also write their pixel count and enter the tile's short root list.
//   2. ccl_seams   pixels on the first row / first column of a tile join the neighbours across the seam in global
//                  memory (paths are two hops long at this point).
//   3. ccl_gather  every local root adds its count to its global root (one atomic per local component, not per
//                  pixel); ccl_argmax: true roots compete for the per-image (area, -root) maximum packed in one u64
//                  atomicMax.  Both walk the per-tile root lists (a few entries per tile) instead of scanning a dense
//                  count array.
//   4. ccl_select  a pixel belongs to the winner iff the root of its local root is the winning root.
// Integer work: bit-exact against the oracle by construction.
#include <stdlib.h>

#include "common.cuh"

namespace wsdl {

constexpr int CT = 32;  // tile edge: one warp per row

__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p = L[i];
  while (p != i) {
    i = p;
    p = L[i];
  }
  return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }  // a > b: hang a under b
    const int old = atomicMin(L + a, b);
    if (old == a) return;
    a = old;  // somebody re-rooted a meanwhile; retry with its new parent
  }
}

// L: -1 background, else global index (within the image) of a pixel of the same component, L[root] == root.
// area: pixel count of the local component at its local root; the other entries are never written NOR read.
// roots: per tile, the number of local roots followed by their global indices (8-connectivity: at most 16 x 16 isolated
// pixels per 32 x 32 tile) -- the later passes walk these short lists instead of scanning `area`.
constexpr int CCL_MAX_ROOTS = (CT / 2) * (CT / 2);
constexpr int CCL_LIST = 1 + CCL_MAX_ROOTS;  // ints per tile

__global__ void __launch_bounds__(CT * CT / 4) ccl_tile(const uint8_t* __restrict__ mask, int* __restrict__ L,
                                                         unsigned* __restrict__ area, int* __restrict__ roots,
                                                         unsigned long long* __restrict__ best, int H, int W) {
  __shared__ int s_lab[CT * CT];
  __shared__ unsigned s_cnt[CT * CT];
  __shared__ int s_roots[CCL_MAX_ROOTS];
  __shared__ int s_nroots;
  if (threadIdx.x == 0) s_nroots = 0;
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
  const size_t img = (size_t)b * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps, 4 rows each
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) best[b] = 0ull;

  // rows: horizontal runs by ballot; a pixel's first label is the first pixel of its run
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, y = y0 + ly, x = x0 + lane;
    const bool fg = y < H && x < W && mask[img + (size_t)y * W + x] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    const unsigned below = ~bits & ((1u << lane) - 1u);            // background pixels left of this one
    const int start = below ? 32 - __clz(below) : 0;               // first pixel after the last of them
    s_lab[ly * CT + lane] = fg ? ly * CT + start : -1;
    s_cnt[ly * CT + lane] = 0u;
  }
  __syncthreads();
  // join with the row above: N, else NW / NE, skipping the joins the left / right neighbour makes anyway
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k;
    if (ly == 0) continue;
    const int i = ly * CT + lane;
    if (s_lab[i] < 0) continue;
    const bool n = s_lab[i - CT] >= 0;
    const bool w = lane > 0 && s_lab[i - 1] >= 0;
    const bool nw = lane > 0 && s_lab[i - CT - 1] >= 0;
    const bool e = lane < CT - 1 && s_lab[i + 1] >= 0;
    const bool ne = lane < CT - 1 && s_lab[i - CT + 1] >= 0;
    if (n) {
      if (!(w && nw)) uf_union(s_lab, i, i - CT);  // else W has joined NW, which is in N's run
    } else {
      if (nw && !w) uf_union(s_lab, i, i - CT - 1);  // else W has joined its N = NW
      if (ne && !e) uf_union(s_lab, i, i - CT + 1);  // else E joins its N = NE
    }
  }
  __syncthreads();
  // flatten, count per local root
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = (warp * 4 + k) * CT + lane;
    if (s_lab[i] < 0) continue;
    const int r = uf_find(s_lab, i);
    s_lab[i] = r;  // roots keep s_lab[r] == r, so concurrent finds stay correct
    atomicAdd(&s_cnt[r], 1u);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, y = y0 + ly, x = x0 + lane;
    if (y >= H || x >= W) continue;
    const int i = ly * CT + lane, r = s_lab[i];
    const size_t g = (size_t)y * W + x;
    L[img + g] = r < 0 ? -1 : (y0 + r / CT) * W + (x0 + r % CT);
    if (r == i) {  // a local root: its count, and an entry in the tile's list
      area[img + g] = s_cnt[i];
      s_roots[atomicAdd(&s_nroots, 1)] = (int)g;
    }
  }
  __syncthreads();
  int* list = roots + ((size_t)(b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * CCL_LIST;
  const int n = s_nroots;
  if (threadIdx.x == 0) list[0] = n;
  for (int k = threadIdx.x; k < n; k += blockDim.x) list[1 + k] = s_roots[k];
}

// joins across the tile seams; one thread per seam pixel (first row and first column of every tile)
__global__ void ccl_seams(int* __restrict__ L, int H, int W) {
  const int b = blockIdx.y;
  int* Lb = L + (size_t)b * H * W;
  const int rows = (H - 1) / CT, cols = (W - 1) / CT;  // interior seams
  const int n_row = rows * W, n_col = cols * H;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_row + n_col; t += gridDim.x * blockDim.x) {
    if (t < n_row) {  // pixel (y, x) in the first row of a tile: N, NW, NE are across the seam
      const int y = (t / W + 1) * CT, x = t % W, i = y * W + x;
      if (Lb[i] < 0) continue;
      const bool n = Lb[i - W] >= 0;
      const bool w = x > 0 && Lb[i - 1] >= 0, nw = x > 0 && Lb[i - W - 1] >= 0;
      const bool e = x + 1 < W && Lb[i + 1] >= 0, ne = x + 1 < W && Lb[i - W + 1] >= 0;
      // the left / right neighbour is in the same tile row only when it is not across a column seam; its own seam
      // join covers ours exactly as inside a tile, whichever tile it is in
      if (n) {
        if (!(w && nw)) uf_union(Lb, i, i - W);
      } else {
        if (nw && !w) uf_union(Lb, i, i - W - 1);
        if (ne && !e) uf_union(Lb, i, i - W + 1);
      }
    } else {  // pixel in the first column of a tile: W, NW, SW are across the seam
      const int u = t - n_row;
      const int x = (u / H + 1) * CT, y = u % H, i = y * W + x;
      if (Lb[i] < 0) continue;
      if (Lb[i - 1] >= 0) uf_union(Lb, i, i - 1);
      if (y > 0 && Lb[i - W - 1] >= 0) uf_union(Lb, i, i - W - 1);
      if (y + 1 < H && Lb[i + W - 1] >= 0) uf_union(Lb, i, i + W - 1);
    }
  }
}

// local roots hand their count to their global root; then (second launch) true roots compete for the maximum
// Local roots hand their count to their global root, then (second launch) true roots compete for the maximum.  Both walk
// the per-tile root lists: eight threads per tile (a tile of a blobby mask holds a handful of local roots).
__global__ void ccl_gather(int* __restrict__ L, unsigned* __restrict__ area, const int* __restrict__ roots, int HW,
                           int tiles_per_image, int n_tiles) {
  const int sub = threadIdx.x & 7;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; t < n_tiles; t += (gridDim.x * blockDim.x) >> 3) {
    const int b = t / tiles_per_image;
    int* Lb = L + (size_t)b * HW;
    unsigned* ab = area + (size_t)b * HW;
    const int* list = roots + (size_t)t * CCL_LIST;
    const int n = list[0];
    for (int k = sub; k < n; k += 8) {
      const int i = list[1 + k];
      const int r = uf_find(Lb, i);
      if (r != i) {
        atomicAdd(ab + r, ab[i]);  // only true roots are added to, and i is not one: ab[i] is still its local count
        Lb[i] = r;                 // compress: pixels of this local component reach the root in two hops
      }
    }
  }
}

__global__ void ccl_argmax(const int* __restrict__ L, const unsigned* __restrict__ area, const int* __restrict__ roots,
                           unsigned long long* __restrict__ best, int HW, int tiles_per_image, int n_tiles) {
  const int sub = threadIdx.x & 7;
  const int stride = (gridDim.x * blockDim.x) >> 3;
  // warp-uniform trip count (the shuffles below need the whole warp): four tiles per warp and iteration
  for (int t0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2; t0 < n_tiles; t0 += stride) {
    const int t = t0 + ((threadIdx.x & 31) >> 3);
    unsigned long long local = 0ull;
    int b = 0;
    if (t < n_tiles) {
      b = t / tiles_per_image;
      const int* Lb = L + (size_t)b * HW;
      const unsigned* ab = area + (size_t)b * HW;
      const int* list = roots + (size_t)t * CCL_LIST;
      const int n = list[0];
      for (int k = sub; k < n; k += 8) {
        const int i = list[1 + k];
        if (Lb[i] == i) {  // a true root
          const unsigned long long key = ((unsigned long long)ab[i] << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
          local = key > local ? key : local;
        }
      }
    }
    // the eight threads of a tile first, then only candidates that beat what the image already holds (the maximum only
    // grows, so a stale read can only let a useless atomic through): hundreds of components per image would otherwise
    // serialise on one address
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, local, o);
      local = other > local ? other : local;
    }
    if (sub == 0 && local && local > *reinterpret_cast<const volatile unsigned long long*>(best + b)) atomicMax(best + b, local);
  }
}

// The select pass reads the labels four pixels per thread and iteration (128-bit loads) when the image size allows
// (VEC: H W % 4 == 0, 4-byte aligned output).
template <bool VEC>
__global__ void ccl_select(const int* __restrict__ L, const unsigned long long* __restrict__ best,
                           uint8_t* __restrict__ out, unsigned* __restrict__ best_area, int HW) {
  const int b = blockIdx.y;
  const unsigned long long key = best[b];
  const int root = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -2;
  const int* Lb = L + (size_t)b * HW;
  uint8_t* ob = out + (size_t)b * HW;
  if (VEC) {
    const int4* L4 = reinterpret_cast<const int4*>(Lb);
    unsigned* o4 = reinterpret_cast<unsigned*>(ob);
    const int nq = HW / 4, step = gridDim.x * blockDim.x;
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int4 nxt = q < nq ? L4[q] : make_int4(-1, -1, -1, -1);
    for (; q < nq; q += step) {
      const int4 p = nxt;
      if (q + step < nq) nxt = L4[q + step];  // the next group's labels travel while this group's roots are looked up
      unsigned v = 0u;
      // neighbours mostly share their local root: one find per distinct value
      const bool w0 = p.x >= 0 && uf_find(Lb, p.x) == root;
      const bool w1 = p.y >= 0 && (p.y == p.x ? w0 : uf_find(Lb, p.y) == root);
      const bool w2 = p.z >= 0 && (p.z == p.y ? w1 : uf_find(Lb, p.z) == root);
      const bool w3 = p.w >= 0 && (p.w == p.z ? w2 : uf_find(Lb, p.w) == root);
      v = (w0 ? 1u : 0u) | (w1 ? 0x100u : 0u) | (w2 ? 0x10000u : 0u) | (w3 ? 0x1000000u : 0u);
      o4[q] = v;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
      const int p = Lb[i];
      ob[i] = (p >= 0 && uf_find(Lb, p) == root) ? 1 : 0;
    }
  }
  if (best_area && blockIdx.x == 0 && threadIdx.x == 0) best_area[b] = (unsigned)(key >> 32);
}

}  // namespace wsdl

using namespace wsdl;

extern "C" size_t wsdl_keep_largest_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t n = (size_t)B * H * W;
  const size_t tiles = (size_t)B * ((H + CT - 1) / CT) * ((W + CT - 1) / CT);
  return 256 + n * 8 + ((size_t)B * 8 + 255) / 256 * 256 + tiles * CCL_LIST * sizeof(int);
}

extern "C" int wsdl_keep_largest(const uint8_t* mask, int B, int H, int W, uint8_t* out, unsigned* best_area,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!mask || !out || !workspace) return WSDL_E_NULL;
  if (B < 1 || H < 1 || W < 1 || B > 65535 || (long long)H * W > 0x7fffffffLL) return WSDL_E_SHAPE;
  if (workspace_bytes < wsdl_keep_largest_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  const int HW = H * W;
  const size_t n = (size_t)B * HW;
  uintptr_t ws = ((uintptr_t)workspace + 255) / 256 * 256;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(ws);
  ws += ((size_t)B * 8 + 255) / 256 * 256;
  int* L = reinterpret_cast<int*>(ws);
  unsigned* area = reinterpret_cast<unsigned*>(ws + n * 4);
  int* roots = reinterpret_cast<int*>(ws + n * 8);
  cudaStream_t s = (cudaStream_t)stream;
  const int tx = (W + CT - 1) / CT, ty = (H + CT - 1) / CT;
  if (ty > 65535) return WSDL_E_SHAPE;
  ccl_tile<<<dim3(tx, ty, B), CT * CT / 4, 0, s>>>(mask, L, area, roots, best, H, W);
  int bx = (HW + 255) / 256;
  // CTAs per SM's worth of grid for the seam / select passes: they are latency bound (dependent label look-ups), and 32
  // waves of 256 threads per SM beat 8 by 17 % at 512^2 (tuning aid: WSDL_CCL_CAP)
  static const int cap_mult = WSDL_TUNE_INT("WSDL_CCL_CAP", 32) > 0 ? WSDL_TUNE_INT("WSDL_CCL_CAP", 32) : 32;
  const int cap = (WSDL_NUM_SMS * cap_mult + B - 1) / B;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, B);
  const int seam_px = ((H - 1) / CT) * W + ((W - 1) / CT) * H;
  if (seam_px > 0) {
    int sb = (seam_px + 255) / 256;
    if (sb > cap) sb = cap < 1 ? 1 : cap;
    ccl_seams<<<dim3(sb, B), 256, 0, s>>>(L, H, W);
  }
  {
    const int tiles_per_image = tx * ty;
    const long long n_tiles = (long long)tiles_per_image * B;
    if (n_tiles > 0x7fffffffLL / 8) return WSDL_E_SHAPE;
    long long gb = (n_tiles * 8 + 255) / 256;
    if (gb > WSDL_NUM_SMS * 16) gb = WSDL_NUM_SMS * 16;
    ccl_gather<<<(int)gb, 256, 0, s>>>(L, area, roots, HW, tiles_per_image, (int)n_tiles);
    ccl_argmax<<<(int)gb, 256, 0, s>>>(L, area, roots, best, HW, tiles_per_image, (int)n_tiles);
  }
  if ((HW & 3) == 0 && ((uintptr_t)out & 3) == 0) {  // four pixels per thread and iteration
    int bv = (HW / 4 + 255) / 256;
    if (bv > cap) bv = cap < 1 ? 1 : cap;
    ccl_select<true><<<dim3(bv, B), 256, 0, s>>>(L, best, out, best_area, HW);
  } else {
    ccl_select<false><<<grid, 256, 0, s>>>(L, best, out, best_area, HW);
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}
