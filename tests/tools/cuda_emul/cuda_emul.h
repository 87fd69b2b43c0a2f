// cuda_emul.h -- TEST INFRASTRUCTURE: a minimal CPU emulation of the CUDA execution model, enough to
// run this repository's SIMT kernels (tensor-core / bulk-copy code: see tc_emul.h; no cluster code) functionally on the host
// when no GPU is available.  One std::thread per CUDA thread, thread blocks run one after another,
// __syncthreads() is a std::barrier, warp shuffles go through a per-warp exchange buffer.  It checks
// indexing, barrier placement and arithmetic order -- not performance, not memory-model races.
// Used only by tests/test_cuda_emul.py (via tests/tools/cuda_emul/build.py); never by the product.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))
#define MSMD_API

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

typedef void* cudaStream_t;
typedef void* msmd_stream_t;
typedef void* cudaEvent_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return 0;
}

#ifndef MSMD_EMUL_WITH_HEADER   // units that paste include/msmd_b200.h get the status codes from it
enum { MSMD_OK = 0, MSMD_ERR_INVALID = -1, MSMD_ERR_CUDA = -2, MSMD_ERR_WORKSPACE = -3 };
#endif
// streams and events: the emulation is synchronous, every launch has completed when it returns
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaMemcpyDeviceToHost = 2, cudaMemcpyHostToDevice = 1 };
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x5; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)0xE; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) {
  memcpy(d, s, n);
  return 0;
}

namespace emu {
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern std::barrier<>* g_block_barrier;
struct WarpBox {
  std::unique_ptr<std::barrier<>> bar;
  long long slot[32];
};
extern std::vector<WarpBox> g_warps;
extern char g_error[512];
extern std::atomic<int> g_or;
// dynamic shared memory + per-CTA hooks (set by tc_emul.h's translation unit; null for the SIMT units)
extern uint8_t* g_dyn_smem;
extern void (*g_block_begin)(uint32_t dyn_smem_bytes);
extern void (*g_block_end)();

template <typename F>
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, F body) {
  // The CTA's threads are created ONCE per launch and walk the grid together (thread blocks run one after
  // another): two launch-level barriers per CTA instead of nthreads thread creations.
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  if (nthreads == 0 || nblocks == 0) return;
  g_blockDim = block;
  g_gridDim = grid;
  std::barrier<> cta_done(nthreads), cta_ready(nthreads);
  std::unique_ptr<std::barrier<>> bar;
  auto begin_cta = [&]() {
    bar.reset(new std::barrier<>(nthreads));
    g_block_barrier = bar.get();
    if (g_block_begin) g_block_begin((uint32_t)dyn_smem_bytes);
    g_warps.clear();
    g_warps.resize((nthreads + 31) / 32);
    for (unsigned w = 0; w < g_warps.size(); ++w) {
      const unsigned lanes = std::min(32u, nthreads - w * 32);
      g_warps[w].bar.reset(new std::barrier<>(lanes));
    }
  };
  begin_cta();
  std::vector<std::thread> ts;
  ts.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t)
    ts.emplace_back([&, t]() {
      t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (unsigned b = 0; b < nblocks; ++b) {
        t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
        body();
        g_block_barrier->arrive_and_drop();            // exited threads leave the CTA's barriers
        g_warps[t / 32].bar->arrive_and_drop();
        cta_done.arrive_and_wait();
        if (t == 0) {                                  // everyone else is parked between the two barriers
          if (g_block_end) g_block_end();
          if (b + 1 < nblocks) begin_cta();
        }
        cta_ready.arrive_and_wait();
      }
    });
  for (auto& th : ts) th.join();
}
template <typename F>
void launch(dim3 grid, dim3 block, F body) { launch(grid, block, 0, body); }
}  // namespace emu

#define threadIdx (::emu::t_threadIdx)
#define blockIdx (::emu::t_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)

static inline void __syncthreads() { ::emu::g_block_barrier->arrive_and_wait(); }

static inline int __syncthreads_or(int pred) {
  if (pred) ::emu::g_or.store(1);
  ::emu::g_block_barrier->arrive_and_wait();
  const int r = ::emu::g_or.load();
  ::emu::g_block_barrier->arrive_and_wait();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) ::emu::g_or.store(0);
  ::emu::g_block_barrier->arrive_and_wait();
  return r;
}
using std::max;
using std::min;

template <typename T>
static inline T emu_warp_exchange(T v, int src_lane_of_me, bool take) {
  static_assert(sizeof(T) <= sizeof(long long), "shuffle payload");
  const unsigned lin = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  auto& box = ::emu::g_warps[lin / 32];
  const int lane = lin & 31;
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  box.slot[lane] = raw;
  box.bar->arrive_and_wait();
  T out = v;
  if (take) memcpy(&out, &box.slot[src_lane_of_me], sizeof(T));
  box.bar->arrive_and_wait();
  return out;
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
  return emu_warp_exchange(v, lane - d, lane - d >= 0);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int d) {
  const int lane = (threadIdx.x + threadIdx.y * blockDim.x) & 31;
  return emu_warp_exchange(v, lane ^ d, true);
}
template <typename T>
static inline unsigned __match_any_sync(unsigned, T v) {
  static_assert(sizeof(T) <= sizeof(long long), "match payload");
  const unsigned lin = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  auto& box = ::emu::g_warps[lin / 32];
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  box.slot[lin & 31] = raw;
  box.bar->arrive_and_wait();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (box.slot[l] == raw ? 1u : 0u) << l;
  box.bar->arrive_and_wait();
  return m;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned m = __match_any_sync(0xffffffffu, pred ? 1 : 0);
  return pred ? m : ~m;
}
template <typename T>
static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline float atomicAdd(float* p, float v) {
  float old = *p, want;
  do { want = old + v; } while (!__atomic_compare_exchange(p, &old, &want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
  return old;
}
template <typename T>
static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <typename T>
static inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <typename T>
static inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  const unsigned lin = threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y;
  ::emu::g_warps[lin / 32].bar->arrive_and_wait();
}
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float2int_rz(float f) { return (int)f; }
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }

// what csrc/common.cuh provides on the host side
namespace msmd {
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(::emu::g_error, sizeof(::emu::g_error), fmt, ap);
  va_end(ap);
}
static inline void count_launch(int) {}
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int kNumSMs = 148;
}  // namespace msmd

#define MSMD_CUDA_OK(expr)                         \
  do {                                             \
    if ((expr) != cudaSuccess) return MSMD_ERR_CUDA; \
  } while (0)
#define MSMD_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::msmd::set_error(__VA_ARGS__);    \
      return MSMD_ERR_INVALID;           \
    }                                    \
  } while (0)
#define MSMD_LAUNCH_OK() \
  do {                   \
  } while (0)
