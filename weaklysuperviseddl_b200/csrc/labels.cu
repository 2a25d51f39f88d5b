// Label assembly and evaluation counters: the two steps either side of the hot path.
//
//  * wsdl_labels_from_masks: what the reference does through a PNG round trip -- save_image of a {0,1} mask
//    (PsuedoMasks.py:67-69 -> {0,255}), reload with convert('L') + NEAREST resize to 256x256 + int64
//    (SegmentationDataset.py:21,26,35) and clamp(max=1) (SegmentationModel.py:100) -- as one kernel on the device
//    masks.  PIL's NEAREST reads source pixel floor((dst + 0.5) * in / out).
//  * wsdl_iou_acc_counts: the three reductions of compute_iou_and_acc (ExtraUtilities.py:4-21: intersection, union,
//    equal pixels), each followed by a host sync in the reference, as one pass and one small read-back.
// Integer work: bit-exact by construction.
#include "common.cuh"

namespace wsdl {

__global__ void labels_from_masks_kernel(const uint8_t* __restrict__ mask, int H, int W, int S, long long* __restrict__ out,
                                         size_t total) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    const size_t r = i / S;
    const int y = (int)(r % S);
    const size_t b = r / S;
    // Pillow (Resample.c / Geometry.c nearest): source = (int)((dst + 0.5) * in / out), in double
    int sx = (int)(((double)x + 0.5) * (double)W / (double)S);
    int sy = (int)(((double)y + 0.5) * (double)H / (double)S);
    sx = min(sx, W - 1), sy = min(sy, H - 1);
    out[i] = mask[(b * H + sy) * W + sx] ? 1 : 0;
  }
}

template <typename T>
__device__ __forceinline__ void iou_item(const void* pred, const void* truth, size_t i, unsigned& inter, unsigned& uni,
                                         unsigned& eq) {
  const T p = reinterpret_cast<const T*>(pred)[i], t = reinterpret_cast<const T*>(truth)[i];
  const bool pf = p > (T)0, tf = t > (T)0;
  inter += (pf && tf), uni += (pf || tf), eq += (p == t);
}

// one CTA row per image (grid.y = B); counts[b] = {intersection, union, equal}
template <typename T>
__global__ void __launch_bounds__(256) iou_acc_kernel(const void* pred, const void* truth, size_t n,
                                                      unsigned long long* counts) {
  const size_t base = (size_t)blockIdx.y * n;
  unsigned inter = 0, uni = 0, eq = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    iou_item<T>(pred, truth, base + i, inter, uni, eq);
  inter = warp_sum_u32(inter), uni = warp_sum_u32(uni), eq = warp_sum_u32(eq);
  if ((threadIdx.x & 31) == 0) {
    if (inter) atomicAdd(counts + 3 * blockIdx.y + 0, (unsigned long long)inter);
    if (uni) atomicAdd(counts + 3 * blockIdx.y + 1, (unsigned long long)uni);
    if (eq) atomicAdd(counts + 3 * blockIdx.y + 2, (unsigned long long)eq);
  }
}

}  // namespace wsdl

using namespace wsdl;

extern "C" int wsdl_labels_from_masks(const uint8_t* mask, int B, int H, int W, int out_size, long long* labels,
                                      void* stream) {
  if (!mask || !labels) return WSDL_E_NULL;
  if (B < 1 || H < 1 || W < 1 || out_size < 1) return WSDL_E_SHAPE;
  if ((uintptr_t)labels % 8) return WSDL_E_ALIGN;
  const size_t total = (size_t)B * out_size * out_size;
  const int threads = 256;
  size_t blocks = (total + threads - 1) / threads;
  if (blocks > (size_t)WSDL_NUM_SMS * 16) blocks = (size_t)WSDL_NUM_SMS * 16;
  labels_from_masks_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(mask, H, W, out_size, labels, total);
  WSDL_LAUNCH_CHECK();
  return 0;
}

// elem: 0 = u8, 1 = i32, 2 = i64, 3 = f32.  counts (B,3) u64 must be zero on entry (they are added to).
extern "C" int wsdl_iou_acc_counts(const void* pred, const void* truth, int B, size_t n, int elem,
                                   unsigned long long* counts, void* stream) {
  if (!pred || !truth || !counts) return WSDL_E_NULL;
  if (B < 1 || B > 65535 || n < 1) return WSDL_E_SHAPE;
  if ((uintptr_t)counts % 8) return WSDL_E_ALIGN;
  size_t bx = (n + 256 * 8 - 1) / (256 * 8);
  const size_t cap = ((size_t)WSDL_NUM_SMS * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  const dim3 grid((unsigned)bx, (unsigned)B);
  cudaStream_t s = (cudaStream_t)stream;
  switch (elem) {
    case 0: iou_acc_kernel<uint8_t><<<grid, 256, 0, s>>>(pred, truth, n, counts); break;
    case 1: iou_acc_kernel<int><<<grid, 256, 0, s>>>(pred, truth, n, counts); break;
    case 2: iou_acc_kernel<long long><<<grid, 256, 0, s>>>(pred, truth, n, counts); break;
    case 3: iou_acc_kernel<float><<<grid, 256, 0, s>>>(pred, truth, n, counts); break;
    default: return WSDL_E_DTYPE;
  }
  WSDL_LAUNCH_CHECK();
  return 0;
}
