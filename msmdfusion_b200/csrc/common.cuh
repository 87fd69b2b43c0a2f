// common.cuh -- shared device/host helpers for libmsmd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/msmd_b200.h"

namespace msmd {

// ---------------------------------------------------------------------------------
// error plumbing: every extern "C" entry point returns 0 or a negative msmd_status
// and leaves a message retrievable through msmd_last_error().
// ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n);  // bumps the counter msmd_launch_count() reports

#define MSMD_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::msmd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                        cudaGetErrorString(_e));                                        \
      return MSMD_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define MSMD_REQUIRE(cond, ...)                                                         \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      ::msmd::set_error(__VA_ARGS__);                                                   \
      return MSMD_ERR_INVALID;                                                          \
    }                                                                                   \
  } while (0)

#define MSMD_LAUNCH_OK()                  \
  do {                                    \
    ::msmd::count_launch(1);              \
    MSMD_CUDA_OK(cudaGetLastError());     \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// Bump allocator over a caller-provided workspace (the library never allocates).
struct Workspace {
  char* base;
  size_t size;
  size_t used;
  Workspace(void* p, size_t n) : base((char*)p), size(n), used(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t off = align_up(used, 256);
    size_t end = off + sizeof(T) * count;
    used = end;
    if (end > size || base == nullptr) return nullptr;
    return (T*)(base + off);
  }
  bool ok() const { return base != nullptr && used <= size; }
};

// ---------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan (blockDim.x a multiple of 32, <= 1024).  `smem` points to
// 33 elements of shared memory.  Returns this thread's exclusive prefix; `total` gets the
// block sum.  Safe to call repeatedly (leading barrier protects smem reuse).
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T& total, T* smem) {
  const T incl = warp_inclusive_scan(v);
  const int w = warp_id(), l = lane_id();
  const int nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    const T s = (l < nw) ? smem[l] : T(0);
    const T si = warp_inclusive_scan(s);
    smem[l] = si - s;
    if (l == 31) smem[32] = si;
  }
  __syncthreads();
  total = smem[32];
  return smem[w] + (incl - v);
}

template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T t = __shfl_xor_sync(0xffffffffu, v, d);
    v = t < v ? t : v;
  }
  return v;
}

}  // namespace msmd
