// Shared device helpers for the wsdl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/wsdl_b200.h"

#define WSDL_NUM_SMS 148

// Tuning knobs are compiled out of the product library: it reads no environment and keeps no mutable state.
// A measurement build (-DWSDL_TUNING, scripts/) turns them into environment look-ups.
#ifdef WSDL_TUNING
#include <stdlib.h>
#define WSDL_TUNE_INT(name, dflt) ([]() { const char* e__ = getenv(name); return e__ ? atoi(e__) : (dflt); }())
#else
#define WSDL_TUNE_INT(name, dflt) (dflt)
#endif

#define WSDL_LAUNCH_CHECK()                    \
  do {                                         \
    cudaError_t e__ = cudaGetLastError();      \
    if (e__ != cudaSuccess) return (int)e__;   \
  } while (0)

namespace wsdl {

// Programmatic dependent launch.  A kernel launched through wsdl::launch_pdl may start while the previous kernel of the
// stream is still draining: it must not touch global memory before pdl_wait() (which returns once that kernel's writes are
// visible), and pdl_launch_dependents() tells the scheduler that ITS successor may be scheduled as soon as every CTA got
// there.  With a predecessor that never signals (any other kernel) the launch simply behaves like a normal one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// 128-bit streaming load: read-only path, do not allocate in L1 (every input byte of the
// channel-sum is touched exactly once).
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned short ld_stream_u16(const void* p) {
  unsigned short r;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
// L2-coherent load (bypasses L1): for data another CTA of the same launch has written.
__device__ __forceinline__ float ld_cg_f32(const float* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned ld_cg_u32(const unsigned* p) { return __ldcg(p); }

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }

// element conversions for 16-bit storage types, from the raw 16-bit pattern
template <int DTYPE>
__device__ __forceinline__ float bits16_to_f32(unsigned short b);
template <>
__device__ __forceinline__ float bits16_to_f32<WSDL_BF16>(unsigned short b) {
  return __uint_as_float(((unsigned)b) << 16);
}
template <>
__device__ __forceinline__ float bits16_to_f32<WSDL_F16>(unsigned short b) {
  return __half2float(__ushort_as_half(b));
}

}  // namespace wsdl
