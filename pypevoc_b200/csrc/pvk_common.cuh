// Shared helpers of libpvk: error reporting, launch wrapper, warp primitives.
#pragma once

#ifdef PVK_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "pvk.h"

namespace pvk {

void set_error(const char *fmt, ...);
void count_launch();

#ifdef PVK_EMU
#define PVK_SMEM(name) unsigned char *name = emu_smem()
#define PVK_LAUNCH(kernel, grid, block, smem, stream, ...)                                 \
  do {                                                                                     \
    auto _body = [=]() { kernel(__VA_ARGS__); };                                           \
    pvk::count_launch();                                                                   \
    emu::launch(_body, grid, block, smem);                                                 \
  } while (0)
#define PVK_SET_SMEM(kernel, bytes) (0)
#else
#define PVK_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PVK_LAUNCH(kernel, grid, block, smem, stream, ...)                                 \
  do {                                                                                     \
    pvk::count_launch();                                                                   \
    kernel<<<grid, block, smem, (cudaStream_t)(stream)>>>(__VA_ARGS__);                    \
  } while (0)
#define PVK_SET_SMEM(kernel, bytes)                                                        \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

#define PVK_CHECK_LAUNCH(what)                                                             \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      pvk::set_error("%s: %s", what, cudaGetErrorString(_e));                              \
      return PVK_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define PVK_REQUIRE(cond, ...)                                                             \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      pvk::set_error(__VA_ARGS__);                                                         \
      return PVK_ERR_ARG;                                                                  \
    }                                                                                      \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

template <class T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ unsigned warp_umin(unsigned v) { return __reduce_min_sync(FULL, v); }   // REDUX
__device__ __forceinline__ unsigned warp_umax(unsigned v) { return __reduce_max_sync(FULL, v); }
// inclusive warp scan (sum)
__device__ __forceinline__ int warp_scan_incl(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL, v, o);
    if (lane_id() >= o) v += t;
  }
  return v;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace pvk
