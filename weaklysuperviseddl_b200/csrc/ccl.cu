// keep_largest: 8-connected component labelling + largest-area selection on the GPU, sm_100a.
//
// Replaces keep_largest (reference TraditionalModel/PsuedoMasks.py:15-21; duplicate at
// AlternatingDirectionCutLoss.py:206-213): skimage.measure.label (full connectivity, background 0),
// regionprops, max by area (first maximum = lowest label = component whose first pixel comes first
// in raster order), `labeled == label`.  Empty masks are returned unchanged.
//
// Union-find with the minimum pixel index as root (so root order == skimage label order):
//   init  -> merge with the 4 already-visited neighbours (W, NW, N, NE) -> flatten + area histogram
//   -> per-image argmax of (area, -root) packed in one u64 atomicMax -> select.
// Integer work: bit-exact against the oracle by construction.
#include "common.cuh"

namespace wsdl {

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

__global__ void ccl_init_local(const uint8_t* __restrict__ mask, int* __restrict__ L, unsigned* __restrict__ area,
                               unsigned long long* __restrict__ best, int HW, int B) {
  const int b = blockIdx.y;
  const uint8_t* m = mask + (size_t)b * HW;
  int* Lb = L + (size_t)b * HW;
  unsigned* ab = area + (size_t)b * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    Lb[i] = m[i] ? i : -1;
    ab[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) best[b] = 0ull;
}

__global__ void ccl_merge(int* __restrict__ L, int H, int W) {
  const int b = blockIdx.y;
  const int HW = H * W;
  int* Lb = L + (size_t)b * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    if (Lb[i] < 0) continue;
    const int y = i / W, x = i - y * W;
    if (x > 0 && Lb[i - 1] >= 0) uf_union(Lb, i, i - 1);
    if (y > 0) {
      if (Lb[i - W] >= 0) {
        uf_union(Lb, i, i - W);
      } else {  // N is background: NW and NE are not yet connected through it
        if (x > 0 && Lb[i - W - 1] >= 0) uf_union(Lb, i, i - W - 1);
        if (x + 1 < W && Lb[i - W + 1] >= 0) uf_union(Lb, i, i - W + 1);
      }
    }
  }
}

__global__ void ccl_flatten_area(int* __restrict__ L, unsigned* __restrict__ area, int HW) {
  const int b = blockIdx.y;
  int* Lb = L + (size_t)b * HW;
  unsigned* ab = area + (size_t)b * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    if (Lb[i] < 0) continue;
    const int r = uf_find(Lb, i);
    Lb[i] = r;  // roots keep L[r] == r, so concurrent finds stay correct
    atomicAdd(ab + r, 1u);
  }
}

__global__ void ccl_argmax(const int* __restrict__ L, const unsigned* __restrict__ area,
                           unsigned long long* __restrict__ best, int HW) {
  const int b = blockIdx.y;
  const int* Lb = L + (size_t)b * HW;
  const unsigned* ab = area + (size_t)b * HW;
  unsigned long long local = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    if (Lb[i] == i) {  // a root
      const unsigned long long key = ((unsigned long long)ab[i] << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
      local = key > local ? key : local;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, local, o);
    local = other > local ? other : local;
  }
  if ((threadIdx.x & 31) == 0 && local) atomicMax(best + b, local);
}

__global__ void ccl_select(const int* __restrict__ L, const unsigned long long* __restrict__ best,
                           uint8_t* __restrict__ out, unsigned* __restrict__ best_area, int HW) {
  const int b = blockIdx.y;
  const unsigned long long key = best[b];
  const int root = key ? (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)) : -2;
  const int* Lb = L + (size_t)b * HW;
  uint8_t* ob = out + (size_t)b * HW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
    ob[i] = (Lb[i] == root) ? 1 : 0;
  if (best_area && blockIdx.x == 0 && threadIdx.x == 0) best_area[b] = (unsigned)(key >> 32);
}

}  // namespace wsdl

using namespace wsdl;

extern "C" size_t wsdl_keep_largest_workspace_bytes(int B, int H, int W) {
  if (B < 1 || H < 1 || W < 1) return 0;
  const size_t n = (size_t)B * H * W;
  return 256 + n * 8 + ((size_t)B * 8 + 255) / 256 * 256;
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
  cudaStream_t s = (cudaStream_t)stream;
  int bx = (HW + 255) / 256;
  const int cap = (WSDL_NUM_SMS * 8 + B - 1) / B;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  dim3 grid(bx, B);
  ccl_init_local<<<grid, 256, 0, s>>>(mask, L, area, best, HW, B);
  ccl_merge<<<grid, 256, 0, s>>>(L, H, W);
  ccl_flatten_area<<<grid, 256, 0, s>>>(L, area, HW);
  ccl_argmax<<<grid, 256, 0, s>>>(L, area, best, HW);
  ccl_select<<<grid, 256, 0, s>>>(L, best, out, best_area, HW);
  WSDL_LAUNCH_CHECK();
  return 0;
}
