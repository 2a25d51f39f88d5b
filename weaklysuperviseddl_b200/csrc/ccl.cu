// keep_largest: 8-connected component labelling + largest-area selection on the GPU, sm_100a.
//
// Replaces keep_largest (reference TraditionalModel/PsuedoMasks.py:15-21; duplicate at
// AlternatingDirectionCutLoss.py:206-213): skimage.measure.label (full connectivity, background 0),
// regionprops, max by area (first maximum = lowest label = component whose first pixel comes first
// in raster order), `labeled == label`.  Empty masks are returned unchanged.
//
// Two-level union-find whose per-PIXEL state is one byte and lives in the output buffer itself:
//   1. ccl_tile    one CTA labels a 32x32 tile in SHARED memory (runs by one ballot per row,
//                  joins to the row above with shared-memory atomicMin unions, flatten, count).  A 32x32 tile has at
//                  most 256 8-connected components, so a pixel is described by ONE BYTE: 0 = background, k + 1 = the
//                  k-th local component of its tile.  That byte is written to `out`; the tile's component table (first
//                  pixel, area, union-find parent = itself) and the labels of its 124 border pixels go to small side
//                  arrays.  (Several tiles per CTA -- WSDL_CCL_TPC -- were measured: the phases are divergent latency
//                  chains that the compiler cannot interleave, and 2 / 4 tiles per CTA cost 5 % / 35 %.)
//   2. ccl_seams   joins across tile seams on the COMPONENT tables (node = tile * 256 + k), reading the compact border
//                  labels: a seam pixel costs two bytes from a dense array instead of a 32-byte sector of a label map.
//   3. ccl_gather  every local component adds its area and its first pixel (min) to its global root;
//      ccl_argmax  true roots compete for the per-image (area, -first pixel) maximum packed in one u64 atomicMax.
//   4. ccl_select  per tile: "does local component k belong to the winner?" for its <= 256 components, then every pixel
//                  turns its byte into 0 / 1 in place.
// HBM traffic per image: mask in, labels out, labels in, result out = 4 bytes per pixel (the round-1 version kept an int32
// label map and a count map: 7x the mask's 2 bytes per pixel; ncu in profiles/).  Integer work: bit-exact against the
// oracle by construction.
#include <stdlib.h>

#include "common.cuh"

namespace wsdl {

constexpr int CT = 32;         // tile edge: one warp per row
#ifndef WSDL_CCL_TPC
#define WSDL_CCL_TPC 1
#endif
constexpr int TPC = WSDL_CCL_TPC;  // tiles per CTA (a 32 x 32 TPC strip)
constexpr int CCL_MAXR = 256;  // components of a 32x32 tile under 8-connectivity: one per aligned 2x2 cell at most
constexpr int CCL_THREADS = 256;
constexpr unsigned short CCL_BG = 0xffffu;

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

struct CclTables {
  int* n_roots;            // [tiles]
  int* first_pix;          // [tiles][256] raster index (within the image) of the component's first pixel; never changes
  int* min_pix;            // [tiles][256] at a global root: first pixel of the whole component
  unsigned* area;          // [tiles][256] local pixel count; at a global root: the component's area
  int* parent;             // [tiles][256] union-find over nodes tile * 256 + k
  unsigned short* border;  // [tiles][4][32]: component index of the top row, bottom row, left column, right column pixels
  unsigned long long* best;  // [B] (area << 32) | (0xffffffff - first pixel)
};

// ncu (profiles/): this kernel is bound by instruction issue (85 % issue-active), and most of its instructions used to be
// per-PIXEL union-find work executed under divergence.  Rows are therefore handled as RUNS: a row is one 32-bit ballot,
// everything a pixel needs to know about its neighbours is bit arithmetic on its row's and the row above's ballots, only
// the first pixel of a run ever walks the union-find, and the other pixels of the run copy from it.
__global__ void __launch_bounds__(CCL_THREADS) ccl_tile(const uint8_t* __restrict__ mask, uint8_t* __restrict__ out,
                                                        CclTables T, int H, int W, int tiles_x, int tiles_y) {
  __shared__ int s_lab[TPC][CT * CT];      // run starts: union-find parent (local index); other pixels: their run start
  __shared__ unsigned s_cnt[TPC][CT * CT];
  __shared__ short s_roots[TPC][CCL_MAXR];
  __shared__ short s_idx[TPC][CT * CT];    // at a root: its component number
  __shared__ unsigned s_bits[TPC][CT];     // foreground ballot of every row
  __shared__ int s_n[TPC];
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * TPC, ty = blockIdx.y;
  const int y0 = ty * CT;
  const size_t img = (size_t)b * H * W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // 8 warps, 4 rows each
  const unsigned lt = (1u << lane) - 1u;
  if (threadIdx.x < TPC) s_n[threadIdx.x] = 0;
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) T.best[b] = 0ull;

  unsigned row[TPC][4];  // this warp's four row ballots
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, y = y0 + ly;
    bool fg[TPC];
#pragma unroll
    for (int u = 0; u < TPC; ++u) {
      const int x = (tx0 + u) * CT + lane;
      fg[u] = y < H && x < W && mask[img + (size_t)y * W + x] != 0;
    }
#pragma unroll
    for (int u = 0; u < TPC; ++u) {
      const unsigned bits = __ballot_sync(0xffffffffu, fg[u]);
      row[u][k] = bits;
      const unsigned below = ~bits & lt;                 // background pixels left of this one
      const int start = below ? 32 - __clz(below) : 0;   // first pixel after the last of them
      s_lab[u][ly * CT + lane] = fg[u] ? ly * CT + start : -1;
      s_cnt[u][ly * CT + lane] = 0u;
      if (lane == 0) s_bits[u][ly] = bits;
    }
  }
  __syncthreads();
  // join with the row above: N, else NW / NE, skipping the joins the left / right neighbour makes anyway.  The unions
  // act on run starts (a pixel stands for the start of its run).
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k;
    if (ly == 0) continue;
#pragma unroll
    for (int u = 0; u < TPC; ++u) {
      const unsigned cur = row[u][k], up = k > 0 ? row[u][k - 1] : s_bits[u][ly - 1];
      if (!((cur >> lane) & 1u)) continue;
      const bool n = (up >> lane) & 1u;
      const bool w = lane > 0 && ((cur >> (lane - 1)) & 1u), nw = lane > 0 && ((up >> (lane - 1)) & 1u);
      const bool e = lane < CT - 1 && ((cur >> (lane + 1)) & 1u), ne = lane < CT - 1 && ((up >> (lane + 1)) & 1u);
      int other = -1;
      if (n) {
        if (!(w && nw)) other = lane;  // else W has joined NW, which is in N's run
      } else {
        if (nw && !w) other = lane - 1;  // else W has joined its N = NW
        if (ne && !e) {
          if (other >= 0) uf_union(s_lab[u], s_lab[u][ly * CT + lane], s_lab[u][(ly - 1) * CT + other]);
          other = lane + 1;  // else E joins its N = NE
        }
      }
      if (other >= 0) uf_union(s_lab[u], s_lab[u][ly * CT + lane], s_lab[u][(ly - 1) * CT + other]);
    }
  }
  __syncthreads();
  // run starts: flatten, add the run's length to the root's count, number the components
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, i = ly * CT + lane;
#pragma unroll
    for (int u = 0; u < TPC; ++u) {
      const unsigned cur = row[u][k];
      const bool start = ((cur >> lane) & 1u) && !(lane > 0 && ((cur >> (lane - 1)) & 1u));
      if (!start) continue;
      const unsigned after = ~cur >> lane;                      // first background pixel at or after this one ends the run
      const int len = after ? __ffs(after) - 1 : CT - lane;
      const int r = uf_find(s_lab[u], i);
      atomicAdd(&s_cnt[u][r], (unsigned)len);
      if (r == i) {
        const int c = atomicAdd(&s_n[u], 1);
        s_roots[u][c] = (short)i;
        s_idx[u][i] = (short)c;
      } else {
        s_lab[u][i] = r;  // roots keep s_lab[r] == r, so concurrent finds stay correct
      }
    }
  }
  __syncthreads();
  // component tables
#pragma unroll
  for (int u = 0; u < TPC; ++u) {
    if (tx0 + u >= tiles_x) continue;
    const int n = s_n[u];
    const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx0 + u;
    if (threadIdx.x == 0) T.n_roots[tile] = n;
    if ((int)threadIdx.x < n) {
      const int r = s_roots[u][threadIdx.x];
      const int g = (y0 + r / CT) * W + ((tx0 + u) * CT + r % CT);
      const size_t node = tile * CCL_MAXR + threadIdx.x;
      T.first_pix[node] = g;
      T.min_pix[node] = g;
      T.area[node] = s_cnt[u][r];
      T.parent[node] = (int)node;
    }
  }
  // one byte per pixel into `out`, two bytes per border pixel into the tile's border record
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, y = y0 + ly;
    const int i = ly * CT + lane;
#pragma unroll
    for (int u = 0; u < TPC; ++u) {
      if (tx0 + u >= tiles_x) continue;
      const int x = (tx0 + u) * CT + lane;
      const int st = s_lab[u][i];                   // -1, my run start (or, at a run start, its root)
      int comp = -1;
      if (st >= 0) {
        const bool start = !(lane > 0 && ((row[u][k] >> (lane - 1)) & 1u));
        const int root = start ? st : s_lab[u][st];  // a run start holds its root (or is one); the others hold their start
        comp = (int)s_idx[u][root];
      }
      if (y < H && x < W) out[img + (size_t)y * W + x] = comp < 0 ? 0 : (uint8_t)(comp == 255 ? 255 : comp + 1);
      const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx0 + u;
      unsigned short* rec = T.border + tile * 128;
      const unsigned short v = comp < 0 ? CCL_BG : (unsigned short)comp;
      if (ly == 0) rec[lane] = v;
      if (ly == CT - 1) rec[32 + lane] = v;
      if (lane == 0) rec[64 + ly] = v;
      if (lane == CT - 1) rec[96 + ly] = v;
    }
  }
}

// component index of a BORDER pixel (y, x) of image b, or CCL_BG
__device__ __forceinline__ int ccl_border(const CclTables& T, int b, int y, int x, int tiles_x, int tiles_y, int* node) {
  const int ty = y / CT, tx = x / CT, ly = y - ty * CT, lx = x - tx * CT;
  const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx;
  const unsigned short* rec = T.border + tile * 128;
  const unsigned short v = ly == 0 ? rec[lx] : (ly == CT - 1 ? rec[32 + lx] : (lx == 0 ? rec[64 + ly] : rec[96 + ly]));
  *node = (int)(tile * CCL_MAXR + v);
  return v;
}

// joins across the tile seams; one thread per seam pixel (first row and first column of every tile)
__global__ void ccl_seams(CclTables T, int H, int W, int tiles_x, int tiles_y) {
  const int b = blockIdx.y;
  const int rows = (H - 1) / CT, cols = (W - 1) / CT;  // interior seams
  const int n_row = rows * W, n_col = cols * H;
  int* P = T.parent;
#define CCL_AT(yy, xx, nd) (ccl_border(T, b, (yy), (xx), tiles_x, tiles_y, &(nd)) != CCL_BG)
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_row + n_col; t += gridDim.x * blockDim.x) {
    int me, a, c, d, dummy;
    if (t < n_row) {  // pixel (y, x) in the first row of a tile: N, NW, NE are across the seam
      const int y = (t / W + 1) * CT, x = t % W;
      if (!CCL_AT(y, x, me)) continue;
      const bool n = CCL_AT(y - 1, x, a);
      const bool w = x > 0 && CCL_AT(y, x - 1, dummy), nw = x > 0 && CCL_AT(y - 1, x - 1, c);
      const bool e = x + 1 < W && CCL_AT(y, x + 1, dummy), ne = x + 1 < W && CCL_AT(y - 1, x + 1, d);
      // the left / right neighbour's own seam join covers ours exactly as inside a tile, whichever tile it is in
      if (n) {
        if (!(w && nw)) uf_union(P, me, a);
      } else {
        if (nw && !w) uf_union(P, me, c);
        if (ne && !e) uf_union(P, me, d);
      }
    } else {  // pixel in the first column of a tile: W, NW, SW are across the seam
      const int u = t - n_row;
      const int x = (u / H + 1) * CT, y = u % H;
      if (!CCL_AT(y, x, me)) continue;
      if (CCL_AT(y, x - 1, a)) uf_union(P, me, a);
      if (y > 0 && CCL_AT(y - 1, x - 1, c)) uf_union(P, me, c);
      if (y + 1 < H && CCL_AT(y + 1, x - 1, d)) uf_union(P, me, d);
    }
  }
#undef CCL_AT
}

// Local components hand their area and first pixel to their global root; eight threads per tile (a tile of a blobby mask
// holds a handful of components).
__global__ void ccl_gather(CclTables T, int n_tiles) {
  const int sub = threadIdx.x & 7;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; t < n_tiles; t += (gridDim.x * blockDim.x) >> 3) {
    const int n = T.n_roots[t];
    for (int k = sub; k < n; k += 8) {
      const int node = t * CCL_MAXR + k;
      const int r = uf_find(T.parent, node);
      if (r != node) {
        atomicAdd(T.area + r, T.area[node]);  // only true roots are added to, and `node` is not one: its entries stay local
        atomicMin(T.min_pix + r, T.first_pix[node]);
        T.parent[node] = r;  // compress
      }
    }
  }
}

__global__ void ccl_argmax(CclTables T, int tiles_per_image, int n_tiles) {
  const int sub = threadIdx.x & 7;
  const int stride = (gridDim.x * blockDim.x) >> 3;
  // warp-uniform trip count (the shuffles below need the whole warp): four tiles per warp and iteration
  for (int t0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2; t0 < n_tiles; t0 += stride) {
    const int t = t0 + ((threadIdx.x & 31) >> 3);
    unsigned long long local = 0ull;
    int b = 0;
    if (t < n_tiles) {
      b = t / tiles_per_image;
      const int n = T.n_roots[t];
      for (int k = sub; k < n; k += 8) {
        const int node = t * CCL_MAXR + k;
        if (T.parent[node] == node) {  // a true root
          const unsigned long long key =
              ((unsigned long long)T.area[node] << 32) | (unsigned long long)(0xffffffffu - (unsigned)T.min_pix[node]);
          local = key > local ? key : local;
        }
      }
    }
    // the eight threads of a tile first, then only candidates that beat what the image already holds (the maximum only
    // grows, so a stale read can only let a useless atomic through)
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, local, o);
      local = other > local ? other : local;
    }
    if (sub == 0 && local && local > *reinterpret_cast<const volatile unsigned long long*>(T.best + b)) atomicMax(T.best + b, local);
  }
}

// One CTA per strip of four tiles: which of each tile's components belong to the winner, then the bytes in place.
constexpr int SPC = 4;  // tiles per CTA of the select pass (straight-line code: wider strips only save CTAs and barriers)
__global__ void __launch_bounds__(CCL_THREADS) ccl_select(uint8_t* __restrict__ out, CclTables T, unsigned* __restrict__ best_area,
                                                          int H, int W, int tiles_x, int tiles_y) {
  __shared__ uint8_t s_flag[SPC][CCL_MAXR];
  __shared__ uint8_t s_cell[SPC][CCL_MAXR];  // only for a tile with 256 components: flag by aligned 2x2 cell
  __shared__ int s_n[SPC];
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * SPC, ty = blockIdx.y, y0 = ty * CT;
  const unsigned long long key = T.best[b];
  const int win = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -2;  // first pixel of the winning component
  if (best_area && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) best_area[b] = (unsigned)(key >> 32);
  {  // 64 threads per tile walk its component table (a handful of entries for a blobby mask)
    const int u = threadIdx.x >> 6, k0 = threadIdx.x & 63;
    int n = 0;
    if (tx0 + u < tiles_x) {
      const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx0 + u;
      n = T.n_roots[tile];
      for (int k = k0; k < n; k += 64) {
        const int node = (int)(tile * CCL_MAXR + k);
        const int r = uf_find(T.parent, node);
        const uint8_t f = T.min_pix[r] == win ? 1 : 0;
        s_flag[u][k] = f;
        if (n == CCL_MAXR) {  // every component sits in its own aligned 2x2 cell: the cell identifies it
          const int g = T.first_pix[node], ly = g / W - y0, lx = g % W - (tx0 + u) * CT;
          s_cell[u][(ly >> 1) * 16 + (lx >> 1)] = f;
        }
      }
    }
    if (k0 == 0) s_n[u] = n;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t img = (size_t)b * H * W;
  const bool vec = ((W & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
  const int x = tx0 * CT + 4 * lane;  // four pixels of tile u = lane / 8
  const int u = lane >> 3;
  if (x >= W) return;
  const bool full = s_n[u] == CCL_MAXR;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ly = warp * 4 + k, y = y0 + ly;
    if (y >= H) continue;
    uint8_t* p = out + img + (size_t)y * W + x;
    if (vec) {
      const unsigned v = *reinterpret_cast<const unsigned*>(p);
      unsigned o = 0u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned l = (v >> (8 * j)) & 0xffu;
        const int lx = (4 * lane + j) & 31;
        const unsigned f = l == 0u ? 0u : (full ? s_cell[u][(ly >> 1) * 16 + (lx >> 1)] : s_flag[u][l - 1u]);
        o |= f << (8 * j);
      }
      *reinterpret_cast<unsigned*>(p) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (x + j >= W) break;
        const unsigned l = p[j];
        const int lx = (4 * lane + j) & 31;
        p[j] = l == 0u ? 0 : (full ? s_cell[u][(ly >> 1) * 16 + (lx >> 1)] : s_flag[u][l - 1u]);
      }
    }
  }
}

}  // namespace wsdl

using namespace wsdl;

static size_t ccl_al(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t wsdl_keep_largest_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t tiles = (size_t)B * ((H + CT - 1) / CT) * ((W + CT - 1) / CT);
  return 256 + ccl_al((size_t)B * 8) + ccl_al(tiles * 4) + 4 * ccl_al(tiles * CCL_MAXR * 4) + ccl_al(tiles * 128 * 2);
}

extern "C" int wsdl_keep_largest(const uint8_t* mask, int B, int H, int W, uint8_t* out, unsigned* best_area,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  if (!mask || !out || !workspace) return WSDL_E_NULL;
  if (B < 1 || H < 1 || W < 1 || B > 65535 || (long long)H * W > 0x7fffffffLL) return WSDL_E_SHAPE;
  if (workspace_bytes < wsdl_keep_largest_workspace_bytes(B, H, W)) return WSDL_E_WORKSPACE;
  const int tx = (W + CT - 1) / CT, ty = (H + CT - 1) / CT;
  const long long n_tiles = (long long)tx * ty * B;
  if (ty > 65535 || n_tiles > 0x7fffffffLL / CCL_MAXR) return WSDL_E_SHAPE;
  const size_t tiles = (size_t)n_tiles;
  uintptr_t ws = ((uintptr_t)workspace + 255) / 256 * 256;
  CclTables T;
  T.best = reinterpret_cast<unsigned long long*>(ws), ws += ccl_al((size_t)B * 8);
  T.n_roots = reinterpret_cast<int*>(ws), ws += ccl_al(tiles * 4);
  T.first_pix = reinterpret_cast<int*>(ws), ws += ccl_al(tiles * CCL_MAXR * 4);
  T.min_pix = reinterpret_cast<int*>(ws), ws += ccl_al(tiles * CCL_MAXR * 4);
  T.area = reinterpret_cast<unsigned*>(ws), ws += ccl_al(tiles * CCL_MAXR * 4);
  T.parent = reinterpret_cast<int*>(ws), ws += ccl_al(tiles * CCL_MAXR * 4);
  T.border = reinterpret_cast<unsigned short*>(ws);
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 strips((tx + TPC - 1) / TPC, ty, B);
  ccl_tile<<<strips, CCL_THREADS, 0, s>>>(mask, out, T, H, W, tx, ty);
  const int seam_px = ((H - 1) / CT) * W + ((W - 1) / CT) * H;
  if (seam_px > 0) {
    // latency bound (dependent look-ups): a few waves of 256-thread CTAs per SM (tuning aid: WSDL_CCL_CAP)
    static const int cap_mult = WSDL_TUNE_INT("WSDL_CCL_CAP", 32) > 0 ? WSDL_TUNE_INT("WSDL_CCL_CAP", 32) : 32;
    const int cap = (WSDL_NUM_SMS * cap_mult + B - 1) / B;
    int sb = (seam_px + 255) / 256;
    if (sb > cap) sb = cap < 1 ? 1 : cap;
    ccl_seams<<<dim3(sb, B), 256, 0, s>>>(T, H, W, tx, ty);
  }
  {
    long long gb = (n_tiles * 8 + 255) / 256;
    if (gb > WSDL_NUM_SMS * 16) gb = WSDL_NUM_SMS * 16;
    ccl_gather<<<(int)gb, 256, 0, s>>>(T, (int)n_tiles);
    ccl_argmax<<<(int)gb, 256, 0, s>>>(T, tx * ty, (int)n_tiles);
  }
  ccl_select<<<dim3((tx + SPC - 1) / SPC, ty, B), CCL_THREADS, 0, s>>>(out, T, best_area, H, W, tx, ty);
  WSDL_LAUNCH_CHECK();
  return 0;
}
