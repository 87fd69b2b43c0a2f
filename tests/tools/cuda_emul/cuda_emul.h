// cuda_emul.h -- TEST INFRASTRUCTURE: a minimal CPU emulation of the CUDA execution model, enough to
// run this repository's SIMT kernels (tensor-core / bulk-copy code: see tc_emul.h; no cluster code) functionally on the host
// when no GPU is available.  One OS thread per warp whose 32 lanes are cooperative fibers, thread blocks run one
// after another, barriers are polling loops that yield, warp shuffles go through a per-warp exchange buffer.  It checks
// indexing, barrier placement and arithmetic order -- not performance, not memory-model races.
// Used only by tests/test_cuda_emul.py (via tests/tools/cuda_emul/build.py); never by the product.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <functional>
#include <mutex>
#include <ucontext.h>
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
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

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
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
template <typename F>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x5; return 0; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)0x5; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)0xE; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) {
  memcpy(d, s, n);
  return 0;
}

namespace emu {
// Execution core: one OS thread per WARP; the 32 lanes of a warp are cooperative fibers (ucontext) that the warp's
// thread schedules round-robin.  Warp-level synchronisation (__syncwarp, shuffles, ballots) never leaves the thread;
// block-level barriers and mbarrier / flag waits are polling loops that yield to the next lane, and to the OS when
// no lane of the warp can make progress.  (The previous core ran one OS thread per CUDA thread: 320 threads per CTA
// of the tensor-core kernels spent most of their time in futex wake-ups.)
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern char g_error[512];
extern std::atomic<int> g_or;
// dynamic shared memory + per-CTA hooks (set by tc_emul.h's translation unit; null for the SIMT units)
extern uint8_t* g_dyn_smem;
extern void (*g_block_begin)(uint32_t dyn_smem_bytes);
extern bool g_blocks_descending;   // order in which emu::launch runs the thread blocks of a grid
extern void (*g_block_end)();

constexpr size_t kLaneStack = 256 * 1024;
struct Lane {
  ucontext_t ctx;
  bool done = false;
  dim3 tidx;
};
struct WarpCtx {
  Lane lanes[32];
  int nlanes = 0, live = 0;
  ucontext_t sched;
  int sync_count = 0;
  unsigned sync_gen = 0;
  long long slot[32];
  int blocked = 0;          // lanes that yielded from a wait since the scheduler last saw progress
  char* stacks = nullptr;   // nlanes x kLaneStack
};
struct BlockSync {
  std::mutex mu;
  int count = 0, live = 0;
  std::atomic<unsigned> gen{0};
};
extern thread_local WarpCtx* t_warp;
extern thread_local int t_lane;
extern BlockSync g_bsync;
extern std::function<void()>* g_body;

static inline void lane_yield(bool blocked) {   // back to the warp's scheduler
  WarpCtx& w = *t_warp;
  if (blocked) ++w.blocked; else w.blocked = 0;
  const int me = t_lane;
  swapcontext(&w.lanes[me].ctx, &w.sched);
}
// a wait loop's body: let the other lanes (and, if the whole warp is waiting, the other warps) run
static inline void blocked_yield() { lane_yield(true); }

static inline void warp_barrier_wait() {
  WarpCtx& w = *t_warp;
  const unsigned gen = w.sync_gen;
  if (++w.sync_count >= w.live) {
    w.sync_count = 0;
    ++w.sync_gen;
    w.blocked = 0;
    return;
  }
  while (w.sync_gen == gen) lane_yield(true);
}
static inline void block_barrier_wait() {
  BlockSync& b = g_bsync;
  unsigned gen;
  {
    std::lock_guard<std::mutex> g(b.mu);
    gen = b.gen.load();
    if (++b.count >= b.live) {
      b.count = 0;
      b.gen.store(gen + 1);
      t_warp->blocked = 0;
      return;
    }
  }
  while (b.gen.load() == gen) lane_yield(true);
}
static inline void lane_exit() {   // an exited thread leaves the CTA's barriers (arrive_and_drop)
  WarpCtx& w = *t_warp;
  w.lanes[t_lane].done = true;
  --w.live;
  if (w.live > 0 && w.sync_count >= w.live) {
    w.sync_count = 0;
    ++w.sync_gen;
  }
  BlockSync& b = g_bsync;
  std::lock_guard<std::mutex> g(b.mu);
  --b.live;
  if (b.live > 0 && b.count >= b.live) {
    b.count = 0;
    b.gen.store(b.gen.load() + 1);
  }
}
static void lane_main() {
  (*g_body)();
  lane_exit();
  t_warp->blocked = 0;
  // returning resumes uc_link = the scheduler
}

template <typename F>
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, F body) {
  const unsigned nthreads = block.x * block.y * block.z;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  if (nthreads == 0 || nblocks == 0) return;
  const unsigned nwarps = (nthreads + 31) / 32;
  g_blockDim = block;
  g_gridDim = grid;
  std::function<void()> fn = body;
  g_body = &fn;
  std::vector<WarpCtx> warps(nwarps);
  for (unsigned w = 0; w < nwarps; ++w) {
    warps[w].nlanes = (int)std::min(32u, nthreads - w * 32);
    warps[w].stacks = (char*)aligned_alloc(4096, kLaneStack * warps[w].nlanes);
  }
  std::barrier<> cta_done(nwarps), cta_ready(nwarps);
  auto begin_cta = [&]() {
    g_bsync.count = 0;
    g_bsync.live = (int)nthreads;
    g_bsync.gen.store(0);
    if (g_block_begin) g_block_begin((uint32_t)dyn_smem_bytes);
  };
  begin_cta();
  std::vector<std::thread> ts;
  ts.reserve(nwarps);
  for (unsigned wi = 0; wi < nwarps; ++wi)
    ts.emplace_back([&, wi]() {
      WarpCtx& w = warps[wi];
      t_warp = &w;
      for (unsigned bi = 0; bi < nblocks; ++bi) {
        // thread blocks run one after another; a persistent kernel whose CTA c consumes what CTAs > c produced
        // first (csrc/spconv_sb.cu) needs them in DESCENDING order -- any order is a legal CUDA schedule
        const unsigned b = g_blocks_descending ? nblocks - 1 - bi : bi;
        t_blockIdx = dim3(b % grid.x, (b / grid.x) % grid.y, b / (grid.x * grid.y));
        w.live = w.nlanes;
        w.sync_count = 0;
        w.blocked = 0;
        for (int l = 0; l < w.nlanes; ++l) {
          Lane& L = w.lanes[l];
          const unsigned t = wi * 32 + (unsigned)l;
          L.done = false;
          L.tidx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          getcontext(&L.ctx);
          L.ctx.uc_stack.ss_sp = w.stacks + (size_t)l * kLaneStack;
          L.ctx.uc_stack.ss_size = kLaneStack;
          L.ctx.uc_link = &w.sched;
          makecontext(&L.ctx, (void (*)())lane_main, 0);
        }
        while (w.live > 0) {
          for (int l = 0; l < w.nlanes; ++l) {
            if (w.lanes[l].done) continue;
            t_lane = l;
            t_threadIdx = w.lanes[l].tidx;
            swapcontext(&w.sched, &w.lanes[l].ctx);
          }
          if (w.live > 0 && w.blocked >= w.live) {   // every live lane is waiting on another warp
            w.blocked = 0;
            std::this_thread::yield();
          }
        }
        cta_done.arrive_and_wait();
        if (wi == 0) {                                 // everyone else is parked between the two barriers
          if (g_block_end) g_block_end();
          if (bi + 1 < nblocks) begin_cta();
        }
        cta_ready.arrive_and_wait();
      }
    });
  for (auto& th : ts) th.join();
  for (auto& w : warps) free(w.stacks);
  g_body = nullptr;
}
template <typename F>
void launch(dim3 grid, dim3 block, F body) { launch(grid, block, 0, body); }
}  // namespace emu

#define threadIdx (::emu::t_threadIdx)
#define blockIdx (::emu::t_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)

static inline void __syncthreads() { ::emu::block_barrier_wait(); }

static inline int __syncthreads_or(int pred) {
  if (pred) ::emu::g_or.store(1);
  ::emu::block_barrier_wait();
  const int r = ::emu::g_or.load();
  ::emu::block_barrier_wait();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) ::emu::g_or.store(0);
  ::emu::block_barrier_wait();
  return r;
}
using std::max;
using std::min;

template <typename T>
static inline T emu_warp_exchange(T v, int src_lane_of_me, bool take) {
  static_assert(sizeof(T) <= sizeof(long long), "shuffle payload");
  auto& box = *::emu::t_warp;
  const int lane = ::emu::t_lane;
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  box.slot[lane] = raw;
  ::emu::warp_barrier_wait();
  T out = v;
  if (take) memcpy(&out, &box.slot[src_lane_of_me], sizeof(T));
  ::emu::warp_barrier_wait();
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
  auto& box = *::emu::t_warp;
  long long raw = 0;
  memcpy(&raw, &v, sizeof(T));
  box.slot[::emu::t_lane] = raw;
  ::emu::warp_barrier_wait();
  unsigned m = 0;
  for (int l = 0; l < box.nlanes; ++l) m |= ((!box.lanes[l].done && box.slot[l] == raw) ? 1u : 0u) << l;
  ::emu::warp_barrier_wait();
  return m;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v) {
  for (int d = 16; d > 0; d >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, d);
  return v;
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
static inline void __syncwarp(unsigned = 0xffffffffu) { ::emu::warp_barrier_wait(); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
template <typename T>
static inline T atomicCAS(T* p, T compare, T val) {
  __atomic_compare_exchange_n(p, &compare, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return compare;   // the old value either way
}
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
