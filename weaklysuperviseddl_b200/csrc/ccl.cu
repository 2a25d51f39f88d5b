// keep_largest: 8-connected component labelling + largest-area selection on the GPU, sm_100a.
//
// Replaces keep_largest (reference TraditionalModel/PsuedoMasks.py:15-21; duplicate at
// AlternatingDirectionCutLoss.py:206-213): skimage.measure.label (full connectivity, background 0),
// regionprops, max by area (first maximum = lowest label = component whose first pixel comes first
// in raster order), `labeled == label`.  Empty masks are returned unchanged.
//
// Two-level union-find on RUNS, rows as 32-bit words.  A 32x32 tile is labelled by ONE WARP, lane = row:
//   1. ccl_tile    the lane packs its row's 32 mask bytes into a word; a run is a maximal group of set bits, a node of the
//                  tile's union-find (shared memory, <= 16 runs per row).  A run touches the runs of the row above under
//                  its dilated span -- one `&` against the upper lane's word (warp shuffle), one union per touched run.
//                  Runs are flattened, counted and numbered: a 32x32 tile has at most 256 8-connected components.  What
//                  leaves the SM is small: the 32 row words, one byte per run (its component number), the tile's component
//                  table (first pixel, area, union-find parent = itself) and the component numbers of its 124 border
//                  pixels.  The mask itself is read once and nothing per-pixel is written.
//   2. ccl_seams   joins across tile seams on the COMPONENT tables (node = tile * 256 + k), one warp per tile again: the
//                  seam with the tile on the left row by row from the two tiles' border records, the seam with the tile
//                  above run by run from the two row words -- no per-pixel label map is ever consulted.
//   3. ccl_gather  every local component adds its area and its first pixel (min) to its global root;
//      ccl_argmax  true roots compete for the per-image (area, -first pixel) maximum packed in one u64 atomicMax.
//   4. ccl_select  one warp per tile again: "does local component k belong to the winner?" for its <= 256 components, then
//                  every lane rebuilds its row from the row word and the run bytes, expands the kept bits to bytes and
//                  stores them (two 128-bit stores per row).
// HBM traffic per image: mask in + result out (2 bytes per pixel, the algorithmic minimum) + ~0.7 KB of tables per tile
// (round 2, first version: one byte of per-pixel state written and read back, 4 bytes per pixel; round 1: an int32 label
// map and a count map, 14).  Integer work: bit-exact against the oracle by construction.
#include <stdlib.h>

#include "common.cuh"

namespace wsdl {

constexpr int CT = 32;          // tile edge: one lane per row, one bit per column
constexpr int CCL_TPC = 8;      // tiles (warps) per CTA: a 256-column strip of one tile row
constexpr int CCL_MAXR = 256;   // components of a 32x32 tile under 8-connectivity: one per aligned 2x2 cell at most
constexpr int CCL_RPR = 16;     // runs of a 32-bit row at most
constexpr int CCL_THREADS = 32 * CCL_TPC;
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

// The tile's union-find lives in shared memory, node = row * 16 + k (raster order, which the root rule needs).  A lane
// mostly touches its own row's nodes: stored at [node] that is a stride of 16 words, a 16-way bank conflict on every
// access (ncu: 6.5 M conflicts for 18 M instructions).  Nodes are therefore STORED at k * 32 + row: lanes land in 32
// different banks; the values (parents) stay node numbers.
__device__ __forceinline__ int ccl_at(int node) { return ((node & (CCL_RPR - 1)) << 5) | (node >> 4); }

__device__ __forceinline__ int uf_find_s(const int* L, int i) {
  int p = L[ccl_at(i)];
  while (p != i) {
    i = p;
    p = L[ccl_at(i)];
  }
  return i;
}

__device__ __forceinline__ void uf_union_s(int* L, int a, int b) {
  while (true) {
    a = uf_find_s(L, a);
    b = uf_find_s(L, b);
    if (a == b) return;
    if (a < b) {
      int t = a;
      a = b;
      b = t;
    }  // a > b: hang a under b
    const int old = atomicMin(L + ccl_at(a), b);
    if (old == a) return;
    a = old;  // somebody re-rooted a meanwhile; retry with its new parent
  }
}

struct CclTables {
  int* n_roots;            // [tiles] components of the tile | (most runs in one of its rows) << 16
  int* first_pix;          // [tiles][256] raster index (within the image) of the component's first pixel; never changes
  int* min_pix;            // [tiles][256] at a global root: first pixel of the whole component
  unsigned* area;          // [tiles][256] local pixel count; at a global root: the component's area
  int* parent;             // [tiles][256] union-find over nodes tile * 256 + k
  unsigned short* border;  // [tiles][4][32]: component number of the top row, bottom row, left column, right column pixels
  unsigned* bits;          // [tiles][32] the row words
  uint8_t* runcomp;        // [tiles][16][32] component number of the k-th run of row r at [k][r]
  unsigned long long* best;  // [B] (area << 32) | (0xffffffff - first pixel)
};

// four mask bytes -> four bits (byte != 0)
__device__ __forceinline__ unsigned ccl_pack4(unsigned v) {
  const unsigned nz = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;  // bit 7 of every non-zero byte
  return ((nz >> 7) * 0x01020408u) >> 24;                                     // gathered into bits 0..3
}

// four bits -> four bytes of 0 / 1
__device__ __forceinline__ unsigned ccl_expand4(unsigned nib) { return (nib * 0x00204081u) & 0x01010101u; }

// the run of `m` that starts at bit a: its length and its bits
__device__ __forceinline__ unsigned ccl_run(unsigned m, int a, int* len) {
  const unsigned after = ~(m >> a);  // bit i: pixel a + i is background (the shift fills with background)
  const int l = __ffs(after) - 1;    // >= 1; after != 0 unless a == 0 and the row is full
  *len = after ? l : CT;
  return (after ? ((1u << l) - 1u) : 0xffffffffu) << a;
}

// number of the run that holds bit x, given the row's run-start bits
__device__ __forceinline__ int ccl_run_of(unsigned starts, int x) { return __popc(starts & ((2u << x) - 1u)) - 1; }

// ncu (profiles/): the per-pixel version of this kernel (one thread per pixel, rows as ballots) issued ~6 warp
// instructions per pixel and was bound by instruction issue.  Here a lane owns a ROW: everything a run needs to know
// about the row above is bit arithmetic on two words, and the loops run over runs (a handful per row for a CAM mask).
__global__ void __launch_bounds__(CCL_THREADS) ccl_tile(const uint8_t* __restrict__ mask, CclTables T, int H, int W, int tiles_x,
                                                        int tiles_y, int vec_ok) {
  __shared__ int s_par[CCL_TPC][CT * CCL_RPR];        // union-find over runs, node = row * 16 + k
  __shared__ unsigned s_cnt[CCL_TPC][CT * CCL_RPR];   // at a root: pixels of the component
  __shared__ unsigned short s_root[CCL_TPC][CCL_MAXR];   // component number -> root node
  __shared__ unsigned short s_first[CCL_TPC][CCL_MAXR];  // component number -> first pixel (row * 32 + column)
  __shared__ uint8_t s_comp[CCL_TPC][CT * CCL_RPR];   // root node -> component number
  __shared__ int s_n[CCL_TPC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x * CCL_TPC + warp;
  pdl_wait();
  pdl_launch_dependents();
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) T.best[b] = 0ull;
  if (tx >= tiles_x) return;  // warps are independent: only __syncwarp below
  const int y0 = ty * CT, x0 = tx * CT, y = y0 + lane;
  const size_t img = (size_t)b * H * W;
  const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx;
  int* par = s_par[warp];
  unsigned* cnt = s_cnt[warp];

  unsigned m = 0u;  // my row
  if (y < H) {
    const uint8_t* p = mask + img + (size_t)y * W + x0;
    if (vec_ok && x0 + CT <= W) {
      const uint4 lo = __ldg(reinterpret_cast<const uint4*>(p)), hi = __ldg(reinterpret_cast<const uint4*>(p) + 1);
      m = ccl_pack4(lo.x) | (ccl_pack4(lo.y) << 4) | (ccl_pack4(lo.z) << 8) | (ccl_pack4(lo.w) << 12) |
          (ccl_pack4(hi.x) << 16) | (ccl_pack4(hi.y) << 20) | (ccl_pack4(hi.z) << 24) | (ccl_pack4(hi.w) << 28);
    } else {
      const int nx = min(CT, W - x0);
      for (int j = 0; j < nx; ++j) m |= (p[j] != 0 ? 1u : 0u) << j;
    }
  }
  // Most tiles of a CAM mask lie entirely outside or entirely inside a blob: their tables are known without any labelling.
  if (!__any_sync(0xffffffffu, m != 0u) || __all_sync(0xffffffffu, m == 0xffffffffu)) {
    const bool full = m != 0u;
    unsigned short* rec = T.border + tile * 128;
    const unsigned short v = full ? (unsigned short)0 : CCL_BG;
    rec[lane] = v, rec[32 + lane] = v, rec[64 + lane] = v, rec[96 + lane] = v;
    T.bits[tile * CT + lane] = m;
    if (full) T.runcomp[tile * CCL_RPR * CT + lane] = 0;
    if (lane == 0) {
      T.n_roots[tile] = full ? (1 | (1 << 16)) : 0;
      if (full) {
        const size_t node = tile * CCL_MAXR;
        T.first_pix[node] = T.min_pix[node] = y0 * W + x0;
        T.area[node] = CT * CT;
        T.parent[node] = (int)node;
      }
    }
    return;
  }
  const unsigned up_raw = __shfl_up_sync(0xffffffffu, m, 1);
  const unsigned up = lane ? up_raw : 0u;  // the row above inside the tile
  const unsigned st = m & ~(m << 1), ust = up & ~(up << 1);  // run starts
  const int nr = __popc(st);
  const int maxruns = __reduce_max_sync(0xffffffffu, nr);
  for (int k = 0; k < nr; ++k) par[k * CT + lane] = lane * CCL_RPR + k, cnt[k * CT + lane] = 0u;  // ccl_at(node)
  if (lane == 0) s_n[warp] = 0;
  __syncwarp();
  {  // A run joins every run of the row above that has a pixel under its span widened by one column either side.  The
     // first such run becomes its parent with a plain store (a node is written by its own lane only; the link goes to
     // a smaller node, like every union) -- the common case costs no find and no atomic.
    unsigned s = st;
    for (int k = 0; s; ++k) {
      const int a = __ffs(s) - 1;
      s &= s - 1u;
      int len;
      const unsigned R = ccl_run(m, a, &len);
      const unsigned touched = (R | (R << 1) | (R >> 1)) & up;
      if (touched) par[k * CT + lane] = (lane - 1) * CCL_RPR + ccl_run_of(ust, __ffs(touched) - 1);
    }
  }
  __syncwarp();
  {  // a run that touches several runs of the row above merges their components: real unions, on the row above
    unsigned s = st;
    for (int k = 0; s; ++k) {
      const int a = __ffs(s) - 1;
      s &= s - 1u;
      int len;
      const unsigned R = ccl_run(m, a, &len);
      const unsigned touched = (R | (R << 1) | (R >> 1)) & up;
      unsigned ts = touched & ~(touched << 1);  // one group of touched bits per run of the row above
      if (ts & (ts - 1u)) {
        const int first = (lane - 1) * CCL_RPR + ccl_run_of(ust, __ffs(ts) - 1);
        ts &= ts - 1u;
        while (ts) {
          const int p = __ffs(ts) - 1;
          ts &= ts - 1u;
          uf_union_s(par, first, (lane - 1) * CCL_RPR + ccl_run_of(ust, p));
        }
      }
    }
  }
  __syncwarp();
  // pointer jumping: a blob that spans the tile is a chain of 32 runs; five halvings leave (almost) every run at its root
#pragma unroll 1
  for (int round = 0; round < 5; ++round) {
    for (int k = 0; k < nr; ++k) {
      const int i = k * CT + lane;
      par[i] = par[ccl_at(par[i])];
    }
    __syncwarp();
  }
  {  // flatten, add the run's length to its root's count, number the components
    unsigned s = st;
    for (int k = 0; s; ++k) {
      const int a = __ffs(s) - 1;
      s &= s - 1u;
      int len;
      ccl_run(m, a, &len);
      const int i = lane * CCL_RPR + k;
      const int r = uf_find_s(par, i);
      atomicAdd(&cnt[ccl_at(r)], (unsigned)len);
      if (r == i) {  // the root is the component's first run in raster order (unions hang the larger node under the smaller)
        const int c = atomicAdd(&s_n[warp], 1);
        s_root[warp][c] = (unsigned short)i;
        s_first[warp][c] = (unsigned short)(lane * CT + a);
        s_comp[warp][ccl_at(i)] = (uint8_t)c;
      } else {
        par[k * CT + lane] = r;  // roots keep par[r] == r, so concurrent finds stay correct
      }
    }
  }
  __syncwarp();
  const int n = s_n[warp];
  if (lane == 0) T.n_roots[tile] = n | (maxruns << 16);
  for (int c = lane; c < n; c += 32) {
    const int i = s_root[warp][c], f = s_first[warp][c];
    const int g = (y0 + (f >> 5)) * W + x0 + (f & 31);
    const size_t node = tile * CCL_MAXR + c;
    T.first_pix[node] = g;
    T.min_pix[node] = g;
    T.area[node] = cnt[ccl_at(i)];
    T.parent[node] = (int)node;
  }
  T.bits[tile * CT + lane] = m;
  for (int k = 0; k < maxruns; ++k)  // par[i] is i's root by now
    T.runcomp[(tile * CCL_RPR + k) * CT + lane] = k < nr ? s_comp[warp][ccl_at(par[k * CT + lane])] : (uint8_t)0;
  // component numbers of the border pixels: top / bottom row (lane = column), left / right column (lane = row)
  unsigned short* rec = T.border + tile * 128;
  {
    const unsigned m0 = __shfl_sync(0xffffffffu, m, 0), st0 = __shfl_sync(0xffffffffu, st, 0);
    const unsigned m31 = __shfl_sync(0xffffffffu, m, CT - 1), st31 = __shfl_sync(0xffffffffu, st, CT - 1);
    rec[lane] = (m0 >> lane) & 1u ? (unsigned short)s_comp[warp][ccl_at(par[ccl_at(ccl_run_of(st0, lane))])] : CCL_BG;
    rec[32 + lane] =
        (m31 >> lane) & 1u ? (unsigned short)s_comp[warp][ccl_at(par[ccl_at((CT - 1) * CCL_RPR + ccl_run_of(st31, lane))])] : CCL_BG;
    rec[64 + lane] = m & 1u ? (unsigned short)s_comp[warp][ccl_at(par[lane])] : CCL_BG;
    rec[96 + lane] = m >> 31 ? (unsigned short)s_comp[warp][ccl_at(par[(nr - 1) * CT + lane])] : CCL_BG;
  }
}

// Joins across the tile seams, on the component tables (node = tile * 256 + component).  One warp per tile:
//   * the seam with the tile on the LEFT, lane = row: my column 0 against its column 31 (rows r-1, r, r+1; the diagonal
//     partners are skipped when the straight one exists -- they are its vertical neighbours, already joined inside the tile);
//   * the seam with the tile ABOVE, lane = run of my row 0: the same run-against-word arithmetic as inside a tile, plus
//     the two corner pixels of the tiles above-left / above-right (every diagonal that crosses a tile corner goes
//     up-left or up-right from some tile's first row, so this covers them all).
__global__ void __launch_bounds__(CCL_THREADS) ccl_seams(CclTables T, int tiles_x, int tiles_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x * CCL_TPC + warp;
  pdl_wait();
  pdl_launch_dependents();
  if (tx >= tiles_x) return;
  const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx;
  int* P = T.parent;
  const unsigned m = T.bits[tile * CT + lane];
  if (!__any_sync(0xffffffffu, m != 0u)) return;  // an empty tile has nothing to join
  // The kernel is a chain of dependent look-ups (ncu: long-scoreboard bound): everything that does not depend on another
  // load is requested here, in one round trip -- the neighbours' row words, the border records, the corner pixels.
  const bool has_l = tx > 0, has_u = ty > 0;
  const size_t tl = tile - 1, tu = tile - tiles_x;
  const unsigned ml = has_l ? T.bits[tl * CT + lane] : 0u;
  const unsigned up = has_u ? T.bits[tu * CT + CT - 1] : 0u;
  const unsigned short my_c0 = T.border[tile * 128 + 64 + lane];                    // component of my column-0 pixel
  const unsigned short l_c31 = has_l ? T.border[tl * 128 + 96 + lane] : CCL_BG;     // left tile, column 31, my row
  const unsigned short c_nw = has_u && tx > 0 ? T.border[(tu - 1) * 128 + 32 + CT - 1] : CCL_BG;
  const unsigned short c_ne = has_u && tx + 1 < tiles_x ? T.border[(tu + 1) * 128 + 32] : CCL_BG;
  const unsigned cur = __shfl_sync(0xffffffffu, m, 0);
  const unsigned st = cur & ~(cur << 1), ust = up & ~(up << 1);
  const bool run_lane = has_u && lane < __popc(st);
  int a = 0, len = 0, me_up = 0, first_up = -1;
  unsigned ts = 0u;
  if (run_lane) {  // second round trip: the component of my run of row 0, and of the first run above it touches
    a = __fns(st, 0, lane + 1);
    const unsigned R = ccl_run(cur, a, &len);
    const unsigned touched = (R | (R << 1) | (R >> 1)) & up;
    ts = touched & ~(touched << 1);
    me_up = (int)(tile * CCL_MAXR) + T.runcomp[(tile * CCL_RPR + lane) * CT];
    if (ts) first_up = (int)(tu * CCL_MAXR) + T.runcomp[(tu * CCL_RPR + ccl_run_of(ust, __ffs(ts) - 1)) * CT + CT - 1];
  }
  if (has_l) {  // my column 0 against the left tile's column 31 (rows r-1, r, r+1): the neighbours' records by shuffle
    const unsigned col = __ballot_sync(0xffffffffu, ml >> 31);
    const unsigned short l_up = (unsigned short)__shfl_up_sync(0xffffffffu, (unsigned)l_c31, 1);
    const unsigned short l_dn = (unsigned short)__shfl_down_sync(0xffffffffu, (unsigned)l_c31, 1);
    if (m & 1u) {
      const int me = (int)(tile * CCL_MAXR) + my_c0, base = (int)(tl * CCL_MAXR);
      if ((col >> lane) & 1u) {
        uf_union(P, me, base + l_c31);
      } else {
        if (lane > 0 && ((col >> (lane - 1)) & 1u)) uf_union(P, me, base + l_up);
        if (lane < CT - 1 && ((col >> (lane + 1)) & 1u)) uf_union(P, me, base + l_dn);
      }
    }
  }
  if (run_lane) {  // my run of row 0 against the runs of the row above
    if (ts) {
      uf_union(P, me_up, first_up);
      ts &= ts - 1u;
      while (ts) {
        const int p = __ffs(ts) - 1;
        ts &= ts - 1u;
        uf_union(P, me_up, (int)(tu * CCL_MAXR) + T.runcomp[(tu * CCL_RPR + ccl_run_of(ust, p)) * CT + CT - 1]);
      }
    }
    // (with a pixel straight above, the corner pixel is its row neighbour)
    if (a == 0 && c_nw != CCL_BG && !(up & 1u)) uf_union(P, me_up, (int)((tu - 1) * CCL_MAXR) + c_nw);
    if (a + len == CT && c_ne != CCL_BG && !(up >> 31)) uf_union(P, me_up, (int)((tu + 1) * CCL_MAXR) + c_ne);
  }
}

// Local components hand their area and first pixel to their global root; eight threads per tile (a tile of a blobby mask
// holds a handful of components).
__global__ void ccl_gather(CclTables T, int n_tiles) {
  pdl_wait();
  pdl_launch_dependents();
  const int sub = threadIdx.x & 7;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; t < n_tiles; t += (gridDim.x * blockDim.x) >> 3) {
    const int n = T.n_roots[t] & 0xffff;
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
  pdl_wait();
  pdl_launch_dependents();
  const int sub = threadIdx.x & 7;
  const int stride = (gridDim.x * blockDim.x) >> 3;
  // warp-uniform trip count (the shuffles below need the whole warp): four tiles per warp and iteration
  for (int t0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 2; t0 < n_tiles; t0 += stride) {
    const int t = t0 + ((threadIdx.x & 31) >> 3);
    unsigned long long local = 0ull;
    int b = 0;
    if (t < n_tiles) {
      b = t / tiles_per_image;
      const int n = T.n_roots[t] & 0xffff;
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

// One warp per tile: flag the tile's components that belong to the winner, then every lane rebuilds its row.
__global__ void __launch_bounds__(CCL_THREADS) ccl_select(uint8_t* __restrict__ out, CclTables T, unsigned* __restrict__ best_area,
                                                          int H, int W, int tiles_x, int tiles_y, int vec_ok) {
  __shared__ uint8_t s_flag[CCL_TPC][CCL_MAXR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x * CCL_TPC + warp;
  pdl_wait();
  pdl_launch_dependents();
  const unsigned long long key = T.best[b];
  const int win = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -2;  // first pixel of the winning component
  if (best_area && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) best_area[b] = (unsigned)(key >> 32);
  if (tx >= tiles_x) return;
  const size_t tile = ((size_t)b * tiles_y + ty) * tiles_x + tx;
  const int nn = T.n_roots[tile], n = nn & 0xffff;
  const unsigned m = T.bits[tile * CT + lane];  // requested with the table look-ups, not after them
  bool mine_any = false, mine_all = true;
  for (int c = lane; c < n; c += 32) {
    const int node = (int)(tile * CCL_MAXR + c);
    const bool f = T.min_pix[uf_find(T.parent, node)] == win;
    s_flag[warp][c] = f ? 1 : 0;
    mine_any |= f, mine_all &= f;
  }
  __syncwarp();
  const bool any = __any_sync(0xffffffffu, mine_any), all = __all_sync(0xffffffffu, mine_all);
  unsigned keep = 0u;  // the bits of my row that belong to the winner
  if (any) {           // (most tiles of a blobby mask hold no part of the winner, or nothing else)
    if (all) {
      keep = m;
    } else {
      unsigned s = m & ~(m << 1);
      for (int k = 0; s; ++k) {
        const int a = __ffs(s) - 1;
        s &= s - 1u;
        int len;
        const unsigned R = ccl_run(m, a, &len);
        if (s_flag[warp][T.runcomp[(tile * CCL_RPR + k) * CT + lane]]) keep |= R;
      }
    }
  }
  const int x0 = tx * CT, y = ty * CT + lane;
  if (y >= H) return;
  uint8_t* p = out + (size_t)b * H * W + (size_t)y * W + x0;
  if (vec_ok && x0 + CT <= W) {
    uint4 lo, hi;
    lo.x = ccl_expand4(keep & 15u), lo.y = ccl_expand4((keep >> 4) & 15u), lo.z = ccl_expand4((keep >> 8) & 15u);
    lo.w = ccl_expand4((keep >> 12) & 15u), hi.x = ccl_expand4((keep >> 16) & 15u), hi.y = ccl_expand4((keep >> 20) & 15u);
    hi.z = ccl_expand4((keep >> 24) & 15u), hi.w = ccl_expand4(keep >> 28);
    reinterpret_cast<uint4*>(p)[0] = lo;
    reinterpret_cast<uint4*>(p)[1] = hi;
  } else {
    const int nx = min(CT, W - x0);
    for (int j = 0; j < nx; ++j) p[j] = (uint8_t)((keep >> j) & 1u);
  }
}

}  // namespace wsdl

using namespace wsdl;

static size_t ccl_al(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t wsdl_keep_largest_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t tiles = (size_t)B * ((H + CT - 1) / CT) * ((W + CT - 1) / CT);
  return 256 + ccl_al((size_t)B * 8) + ccl_al(tiles * 4) + 4 * ccl_al(tiles * CCL_MAXR * 4) + ccl_al(tiles * 128 * 2) +
         ccl_al(tiles * CT * 4) + ccl_al(tiles * CCL_RPR * CT);
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
  T.border = reinterpret_cast<unsigned short*>(ws), ws += ccl_al(tiles * 128 * 2);
  T.bits = reinterpret_cast<unsigned*>(ws), ws += ccl_al(tiles * CT * 4);
  T.runcomp = reinterpret_cast<uint8_t*>(ws);
  cudaStream_t s = (cudaStream_t)stream;
  const dim3 strips((tx + CCL_TPC - 1) / CCL_TPC, ty, B);
  // 128-bit row accesses: rows of a tile start on 16-byte boundaries
  const int vec_in = (W % 16 == 0) && ((uintptr_t)mask % 16 == 0), vec_out = (W % 16 == 0) && ((uintptr_t)out % 16 == 0);
  // five short dependent kernels: each is launched so that its CTAs are scheduled under the tail of the one before
  cudaError_t e = launch_pdl(ccl_tile, strips, dim3(CCL_THREADS), 0, s, mask, T, H, W, tx, ty, vec_in);
  if (e != cudaSuccess) return (int)e;
  if (tx > 1 || ty > 1) e = launch_pdl(ccl_seams, strips, dim3(CCL_THREADS), 0, s, T, tx, ty);
  if (e != cudaSuccess) return (int)e;
  {
    long long gb = (n_tiles * 8 + 255) / 256;
    if (gb > WSDL_NUM_SMS * 16) gb = WSDL_NUM_SMS * 16;
    e = launch_pdl(ccl_gather, dim3((unsigned)gb), dim3(256), 0, s, T, (int)n_tiles);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(ccl_argmax, dim3((unsigned)gb), dim3(256), 0, s, T, tx * ty, (int)n_tiles);
    if (e != cudaSuccess) return (int)e;
  }
  e = launch_pdl(ccl_select, strips, dim3(CCL_THREADS), 0, s, out, T, best_area, H, W, tx, ty, vec_out);
  if (e != cudaSuccess) return (int)e;
  WSDL_LAUNCH_CHECK();
  return 0;
}
