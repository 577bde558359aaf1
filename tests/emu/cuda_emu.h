// TEST INFRASTRUCTURE ONLY -- a tiny SIMT emulator so the .cu sources of libpvk can be
// compiled with g++ and their *logic* debugged on the GPU-less build container.
//
// Each CUDA thread of a block is a ucontext fiber; __syncthreads / warp collectives are
// cooperative barriers (round-robin yield).  Blocks run sequentially per OS thread and in
// parallel across OS threads.  Deterministic, so it finds logic / indexing / divergent-
// barrier bugs, not data races.  Never shipped, never loaded by the product
// (pypevoc_b200/_lib.py only loads the nvcc-built libpvk.so and fails loudly without it).
#pragma once
#include <ucontext.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <math.h>
#include <vector>
#include <thread>
#include <functional>
#include <atomic>
#include <algorithm>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3 { unsigned x, y, z; };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

typedef void *cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

namespace emu {
struct Warp { int count; unsigned gen; uint64_t scratch[32]; };
struct Block {
  int nthreads = 0;
  std::vector<ucontext_t> ctx;
  ucontext_t main;
  std::vector<char> done;
  std::vector<char *> stacks;
  int cur = 0;
  int bar_count = 0; unsigned bar_gen = 0;
  std::vector<Warp> warps;
  unsigned char *smem = nullptr;
  std::function<void()> body;
  long progress = 0;
};
extern thread_local Block *g_blk;
void yield();
void run_block(Block &b, dim3 bid, dim3 bdim, dim3 gdim);
void launch(std::function<void()> body, dim3 grid, dim3 block, size_t smem);
}  // namespace emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

static inline unsigned char *emu_smem() { return emu::g_blk->smem; }

static inline void __syncthreads() {
  emu::Block *b = emu::g_blk;
  unsigned gen = b->bar_gen;
  b->progress++;
  if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen++; }
  else while (b->bar_gen == gen) emu::yield();
}
static inline void emu_warp_barrier() {
  emu::Block *b = emu::g_blk;
  emu::Warp &w = b->warps[threadIdx.x >> 5];
  unsigned gen = w.gen;
  b->progress++;
  if (++w.count == 32) { w.count = 0; w.gen++; }
  else while (w.gen == gen) emu::yield();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier(); }

template <class T> static inline uint64_t emu_bits(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_unbits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T> static inline T emu_exchange(T v, int src, bool valid) {
  static_assert(sizeof(T) <= 8, "");
  emu::Warp &w = emu::g_blk->warps[threadIdx.x >> 5];
  w.scratch[threadIdx.x & 31] = emu_bits(v);
  emu_warp_barrier();
  T r = valid ? emu_unbits<T>(w.scratch[src & 31]) : v;
  emu_warp_barrier();
  return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src, true); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) { int l = threadIdx.x & 31; return emu_exchange(v, l - (int)d, l >= (int)d); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) { int l = threadIdx.x & 31; return emu_exchange(v, l + (int)d, l + (int)d < 32); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { int l = threadIdx.x & 31; return emu_exchange(v, l ^ m, true); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::Warp &w = emu::g_blk->warps[threadIdx.x >> 5];
  w.scratch[threadIdx.x & 31] = pred ? 1 : 0;
  emu_warp_barrier();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) if (w.scratch[i]) m |= 1u << i;
  emu_warp_barrier();
  return m;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  emu::Warp &w = emu::g_blk->warps[threadIdx.x >> 5];
  w.scratch[threadIdx.x & 31] = v;
  emu_warp_barrier();
  unsigned m = 0xffffffffu;
  for (int i = 0; i < 32; ++i) if ((unsigned)w.scratch[i] < m) m = (unsigned)w.scratch[i];
  emu_warp_barrier();
  return m;
}
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::Warp &w = emu::g_blk->warps[threadIdx.x >> 5];
  w.scratch[threadIdx.x & 31] = v;
  emu_warp_barrier();
  unsigned m = 0u;
  for (int i = 0; i < 32; ++i) if ((unsigned)w.scratch[i] > m) m = (unsigned)w.scratch[i];
  emu_warp_barrier();
  return m;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline int __syncthreads_count(int p) {
  // block-wide count via warp ballots + shared scratch in warp structs
  unsigned m = __ballot_sync(0xffffffffu, p);
  emu::Block *b = emu::g_blk;
  if ((threadIdx.x & 31) == 0) b->warps[threadIdx.x >> 5].scratch[0] = __builtin_popcount(m);
  __syncthreads();
  int tot = 0;
  for (size_t i = 0; i < b->warps.size(); ++i) tot += (int)b->warps[i].scratch[0];
  __syncthreads();
  return tot;
}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline long long __double_as_longlong(double d) { long long u; memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(long long u) { double d; memcpy(&d, &u, 8); return d; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdividef(float a, float b) { return a / b; }
#define __cosf(a) cosf(a)
#define __sinf(a) sinf(a)
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline double cospi(double a) { return cos(M_PI * a); }
static inline float cospif(float a) { return (float)cos(M_PI * (double)a); }
static inline void sincospi(double a, double *s, double *c) { *s = sin(M_PI * a); *c = cos(M_PI * a); }
static inline void sincospif(float a, float *s, float *c) { *s = (float)sin(M_PI * (double)a); *c = (float)cos(M_PI * (double)a); }
static inline float __double2float_rn(double d) { return (float)d; }
static inline int __double2int_rn(double d) { return (int)rint(d); }
static inline long long __double2ll_rd(double d) { return (long long)floor(d); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
static inline float fminf_(float a, float b) { return fminf(a, b); }

template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float *p, float v) { float o = *p; *p = o + v; return o; }   // single OS thread per block only
static inline double atomicAdd(double *p, double v) { double o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; while (o > v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {} return o; }
template <class T> static inline T atomicOr(T *p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicCAS(T *p, T c, T v) { __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return c; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
#define cudaMemcpyHostToDevice 1
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return 0; }
