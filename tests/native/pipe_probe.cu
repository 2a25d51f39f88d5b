// Stand-alone throughput probe for the FP32 pipes of sm_100a (FFMA vs FFMA2/FADD2/FMUL2, MUFU.EX2).
// Debugging / design aid for pairwise.cu, not a test.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 4096;
constexpr int NACC = 8;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) {  // 2 scalar FFMA
        acc[i].x = fmaf(acc[i].x, a2.x, b2.x);
        acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
      } else if (MODE == 1) {  // 1 FFMA2
        acc[i] = __ffma2_rn(acc[i], a2, b2);
      } else if (MODE == 2) {  // 1 FADD2
        acc[i] = __fadd2_rn(acc[i], b2);
      } else if (MODE == 3) {  // 1 FMUL2
        acc[i] = __fmul2_rn(acc[i], a2);
      } else if (MODE == 4) {  // 2 scalar FADD
        acc[i].x = acc[i].x + b2.x;
        acc[i].y = acc[i].y + b2.y;
      } else if (MODE == 5) {  // 2 MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].y));
      } else if (MODE == 6) {  // mix: 1 MUFU + 7 FFMA2  (the shape of a packed pair evaluation)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].x));
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i] = __ffma2_rn(acc[i], a2, b2);
      } else if (MODE == 7) {  // mix: 1 MUFU + 14 FFMA (the scalar equivalent)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].x));
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          acc[i].x = fmaf(acc[i].x, a2.x, b2.x);
          acc[i].y = fmaf(acc[i].y, a2.y, b2.y);
        }
      } else if (MODE == 8) {  // 2 MUFU + 7 FFMA2 (two pixels' exps, packed math)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].x));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(acc[i].y));
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[i] = __ffma2_rn(acc[i], a2, b2);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, double flop_lanes_per_iter_acc, double mufu_per_iter_acc) {
  float* out;
  const int blocks = 148 * 8;
  cudaMalloc(&out, blocks * 256 * 4);
  k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<blocks, 256>>>(out, 1.0001f, 1e-4f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)blocks * 256 * ITERS * NACC;
  printf("%-28s %8.3f ms  fp32-lane-ops %7.2f T/s   mufu %6.2f T/s  (%s)\n", name, ms, n * flop_lanes_per_iter_acc / ms * 1e-9,
         n * mufu_per_iter_acc / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() {
  run<0>("2x FFMA", 2, 0);
  run<1>("1x FFMA2", 2, 0);
  run<2>("1x FADD2", 2, 0);
  run<3>("1x FMUL2", 2, 0);
  run<4>("2x FADD", 2, 0);
  run<5>("2x MUFU.EX2", 0, 2);
  run<6>("1 MUFU + 7 FFMA2", 14, 1);
  run<7>("1 MUFU + 14 FFMA", 14, 1);
  run<8>("2 MUFU + 7 FFMA2", 14, 2);
  return 0;
}
