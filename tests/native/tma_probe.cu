// Stand-alone probe of the 4-D TMA tile load used by pairwise_tma_kernel (debugging aid, not a test).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int bx, int by, int c, int x0, int y0, int bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(sm)), "l"(&tm), "r"(x0), "r"(y0), "r"(0), "r"(0), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  const float* f = reinterpret_cast<const float*>(sm);
  for (int i = threadIdx.x; i < bx * by * c; i += blockDim.x) out[i] = f[i];
}
int main(int argc, char** argv) {
  int W = atoi(argv[1]), H = atoi(argv[2]), C = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]), x0 = atoi(argv[6]), y0 = atoi(argv[7]);
  int B = 2;
  std::vector<float> h((size_t)W * H * C * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, (size_t)bx * by * C * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)C, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d q=%d\n", (int)r, (int)q);
  int bytes = bx * by * C * 4;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  probe<<<1, 128, bytes + 128>>>(tm, out, bx, by, C, x0, y0, bytes);
  cudaError_t e = cudaDeviceSynchronize();
  printf("W=%d H=%d C=%d box=%dx%d at (%d,%d): %s\n", W, H, C, bx, by, x0, y0, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> o((size_t)bx * by * C);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < C; ++c) for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
      int gx = x0 + x, gy = y0 + y;
      float want = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : (float)(((size_t)c * H + gy) * W + gx);
      if (o[((size_t)c * by + y) * bx + x] != want) ++bad;
    }
    printf("mismatches: %d\n", bad);
  }
  return 0;
}
