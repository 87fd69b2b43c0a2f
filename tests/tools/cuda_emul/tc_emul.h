// tc_emul.h -- TEST INFRASTRUCTURE: a functional host model of what csrc/tc.cuh wraps (mbarrier, the bulk
// copy engine, tensor memory, tcgen05.mma / commit / ld / st, UMMA descriptors), so that the warp-
// specialised tensor-core kernels of csrc/spconv_tc.cu run on the CPU emulator (cuda_emul.h) as they
// are.  It replaces tc.cuh textually (build.py) and keeps its names and signatures.
//
// What it models                               what it does NOT model
//   shared-memory window with real 32-bit        proxy fences, memory-model races between ordinary
//   shared addresses, 1024-byte alignment        threads, timing, bank conflicts, TMEM allocation
//   rules, SWIZZLE_128B K-major operand reads    contention between CTAs.  Asynchrony of the tensor core /
//   (address-bit XOR), descriptor / idesc        copy engine IS modelled, adversarially: see AsyncOp below
//   field decoding with validity checks,
//   mbarrier phases + transaction bytes,
//   TMEM as 128 lanes x 512 columns with the
//   warp -> lane-quarter access rule, tf32
//   operand truncation, fp32 accumulation
//
// Calibration: the kernels it runs are verified on the B200 by tests/test_gpu_parity.py; the emulator
// must reproduce the oracle on them (tests/test_cuda_emul.py::test_emulator_calibration_tc_*) before
// any un-run variant is trusted to it.  A protocol deadlock shows up as a 30 s wait and aborts the process
// with a message naming the barrier.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <map>
#include <mutex>

#include "cuda_emul.h"

namespace emu {
constexpr uint32_t kSmemWindow = 1u << 18;   // shared-address space of one CTA (18 address bits)
constexpr uint32_t kDynBase = 1024 + 16;     // dynamic smem starts 16-byte aligned only, as on the device
extern uint8_t* g_smem_window;               // 2^18-aligned host block: shared address a <-> window + a
extern uint32_t g_dyn_bytes;
extern std::mutex g_tc_mu;
extern std::condition_variable g_tc_cv;
struct MBar {
  uint32_t expected = 0;
  int pending = 0;
  long long tx = 0;
  uint32_t phase = 0;
  bool init = false;
};
extern std::map<uint32_t, MBar> g_mbar;      // keyed by shared address
extern uint32_t g_tmem[128][512];
extern uint32_t g_tmem_next, g_tmem_live;
struct NamedBar {
  int count = 0;
  unsigned gen = 0;
};
extern NamedBar g_named[16];
extern std::atomic<long long> g_mma_count;
// Asynchronous units, scheduled ADVERSARIALLY: a tcgen05.mma / bulk copy is queued at issue and executes as
// LATE as the program allows -- when some thread first polls the mbarrier its completion is tied to (the
// commit's barrier / the copy's complete_tx barrier), or at CTA exit.  Operands are therefore read, and
// results written, at the last legal moment: a kernel that overwrites a stage before waiting on its
// "empty" barrier, or reads a stage / accumulator before waiting on its "full" barrier, computes garbage.
// (g_async_late = 0 restores completion at issue.)
struct AsyncOp {
  std::function<void()> run;   // empty for a commit marker
  uint32_t bar = 0;            // commit marker / bulk copy: shared address of the mbarrier
  uint32_t tx = 0;             // bulk copy: bytes to complete
  int kind = 0;                // 0 = MMA, 1 = commit marker, 2 = bulk copy
};
extern std::deque<AsyncOp> g_mma_queue;        // in issue order (one issuing thread per CTA)
extern std::vector<AsyncOp> g_copy_queue;      // bulk copies complete independently of each other
extern std::map<int, std::vector<std::function<void()>>> g_cpasync_pending;   // per thread: cp.async not yet tied to a barrier
extern int g_async_late;

[[noreturn]] static inline void tc_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  fprintf(stderr, "[tc_emul] FATAL: ");
  vfprintf(stderr, fmt, ap);
  fprintf(stderr, "\n");
  va_end(ap);
  fflush(stderr);
  abort();
}

static inline uint8_t* smem_ptr(uint32_t addr, uint32_t bytes) {
  if (addr < kDynBase || addr + bytes > kDynBase + g_dyn_bytes)
    tc_fail("shared-memory access [%u, +%u) outside the dynamic allocation [%u, +%u)", addr, bytes, kDynBase,
            g_dyn_bytes);
  return g_smem_window + addr;
}

static inline void tc_block_reset(uint32_t dyn_bytes) {
  if (!g_smem_window) g_smem_window = (uint8_t*)aligned_alloc(kSmemWindow, kSmemWindow);
  if (dyn_bytes > 227u * 1024u) tc_fail("dynamic shared memory %u > 227 KB", dyn_bytes);
  g_dyn_bytes = dyn_bytes;
  memset(g_smem_window, 0xEE, kSmemWindow);              // canary outside the allocation
  memset(g_smem_window + kDynBase, 0xCD, dyn_bytes);     // shared memory starts uninitialised
  g_mbar.clear();
  memset(g_tmem, 0xFF, sizeof(g_tmem));                  // NaN pattern: unwritten accumulators show up
  g_tmem_next = 0;
  g_tmem_live = 0;
  for (auto& b : g_named) b = NamedBar();
  g_mma_queue.clear();
  g_copy_queue.clear();
  g_cpasync_pending.clear();
}
static inline void tc_block_check() {
  // whatever is still queued completes now (nobody looked at its barrier any more); the arrivals
  // they would perform cannot matter after the CTA's exit
  for (auto& op : g_mma_queue) if (op.kind == 0) op.run();
  for (auto& op : g_copy_queue) op.run();
  g_mma_queue.clear();
  g_copy_queue.clear();
  if (g_tmem_live != 0) tc_fail("CTA exited with %u TMEM columns still allocated", g_tmem_live);
  for (uint32_t a = 0; a < kSmemWindow; ++a) {
    if (a >= kDynBase && a < kDynBase + g_dyn_bytes) { a = kDynBase + g_dyn_bytes - 1; continue; }
    if (g_smem_window[a] != 0xEE) tc_fail("shared-memory write outside the allocation at address %u", a);
  }
}
static inline int emu_lin_tid() {
  return (int)(threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y);
}
}  // namespace emu

static inline long long clock64() {
  // every bounded spin loop of this code base reads the clock in its body: the one place where a raw polling loop
  // of kernel code hands control to the other lanes of the cooperative warp
  if (::emu::t_warp) ::emu::blocked_yield();
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline void __trap() { ::emu::tc_fail("__trap() reached"); }
template <typename T>
static inline void __stcg(T* p, T v) { *p = v; }
template <typename T>
static inline T __ldcg(const T* p) { return *p; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }

namespace msmd {
namespace tc {

static inline uint32_t smem_u32(const void* p) {
  const uintptr_t d = (uintptr_t)p - (uintptr_t)::emu::g_smem_window;
  if (d >= ::emu::kSmemWindow) ::emu::tc_fail("smem_u32: pointer is not in shared memory");
  return (uint32_t)d;
}

static inline bool elect_one() {   // a converged warp elects one lane: the model always picks lane 0
  __syncwarp();
  return (::emu::emu_lin_tid() & 31) == 0;
}

// ---- mbarrier -------------------------------------------------------------------------
static inline ::emu::MBar& bar_at(uint64_t* bar, bool must_exist = true) {
  const uint32_t a = smem_u32(bar);
  if (a & 7) ::emu::tc_fail("mbarrier at %u is not 8-byte aligned", a);
  ::emu::smem_ptr(a, 8);
  auto& b = ::emu::g_mbar[a];
  if (must_exist && !b.init) ::emu::tc_fail("mbarrier at %u used before mbarrier.init", a);
  return b;
}
static inline void bar_check(::emu::MBar& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = (int)b.expected;
  }
}
static inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar, false);
  b = ::emu::MBar();
  b.expected = count;
  b.pending = (int)count;
  b.init = true;
}
static inline void fence_mbar_init() {}
static inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar);
  if (b.pending <= 0) ::emu::tc_fail("mbarrier at %u: more arrivals than its count %u", smem_u32(bar), b.expected);
  --b.pending;
  bar_check(b);
}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar);
  if (b.pending <= 0) ::emu::tc_fail("mbarrier at %u: more arrivals than its count %u", smem_u32(bar), b.expected);
  b.tx += bytes;
  --b.pending;
  bar_check(b);
}
// g_tc_mu held.  Somebody is about to look at the barrier at shared address `a`: everything whose
// completion that barrier tracks completes now (MMAs in issue order up to the last commit on it).
static inline void async_flush_for(uint32_t a) {
  int last = -1;
  for (int i = 0; i < (int)::emu::g_mma_queue.size(); ++i)
    if (::emu::g_mma_queue[i].kind == 1 && ::emu::g_mma_queue[i].bar == a) last = i;
  for (int i = 0; i <= last; ++i) {
    ::emu::AsyncOp op = std::move(::emu::g_mma_queue.front());
    ::emu::g_mma_queue.pop_front();
    if (op.kind == 0) { op.run(); continue; }
    auto& b = ::emu::g_mbar[op.bar];
    if (b.pending <= 0) ::emu::tc_fail("mbarrier at %u: more arrivals than its count %u", op.bar, b.expected);
    --b.pending;
    bar_check(b);
  }
  for (size_t i = 0; i < ::emu::g_copy_queue.size();) {
    if (::emu::g_copy_queue[i].bar != a) { ++i; continue; }
    ::emu::AsyncOp op = std::move(::emu::g_copy_queue[i]);
    ::emu::g_copy_queue.erase(::emu::g_copy_queue.begin() + (long)i);
    op.run();
    auto& b = ::emu::g_mbar[op.bar];
    if (op.kind == 3) {   // a thread's cp.async group: its (pre-counted) arrival
      if (b.pending <= 0) ::emu::tc_fail("mbarrier at %u: more arrivals than its count %u", op.bar, b.expected);
      --b.pending;
    } else {
      b.tx -= op.tx;
    }
    bar_check(b);
  }
}
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar);
  async_flush_for(smem_u32(bar));
  return (b.phase & 1u) != (parity & 1u);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  const auto t0 = std::chrono::steady_clock::now();
  for (long spin = 0;; ++spin) {
    if (mbar_try_wait(bar, parity)) return;
    ::emu::blocked_yield();
    if ((spin & 0xFFF) == 0xFFF && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) {
      std::lock_guard<std::mutex> g(::emu::g_tc_mu);
      auto& b = bar_at(bar);
      ::emu::tc_fail("deadlock: thread %d waited 30 s on the mbarrier at %u (parity %u, pending %d, tx %lld)",
                     ::emu::emu_lin_tid(), smem_u32(bar), parity, b.pending, b.tx);
    }
  }
}
static inline void mbar_wait_long(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
static inline void griddep_launch_dependents() {}
static inline void griddep_wait() {}
static inline void fence_proxy_async() {}

static inline void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t d = smem_u32(dst_smem);
  if ((d & 15) || ((uintptr_t)src & 15) || (bytes & 15))
    ::emu::tc_fail("cp.async.bulk: dst %u / src %p / size %u must be 16-byte aligned", d, src, bytes);
  uint8_t* dst = ::emu::smem_ptr(d, bytes);
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar);
  if (!::emu::g_async_late) {
    memcpy(dst, src, bytes);
    b.tx -= bytes;
    bar_check(b);
    return;
  }
  ::emu::AsyncOp op;
  op.kind = 2;
  op.bar = smem_u32(bar);
  op.tx = bytes;
  op.run = [dst, src, bytes] { memcpy(dst, src, bytes); };
  ::emu::g_copy_queue.push_back(std::move(op));
}

// cp.async (LDGSTS) with mbarrier completion: the copies a thread issued are parked per thread and run -- as late as
// the program allows, like every asynchronous unit here -- when somebody polls the mbarrier that the thread's
// cp.async.mbarrier.arrive.noinc tied them to; that barrier then receives the thread's arrival
static inline void cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  if ((dst_smem & 15) || ((uintptr_t)src & 15) || (src_bytes != 0 && src_bytes != 16))
    ::emu::tc_fail("cp.async: dst %u / src %p must be 16-byte aligned, src-size %u must be 0 or 16", dst_smem, src,
                   src_bytes);
  uint8_t* dst = ::emu::smem_ptr(dst_smem, 16);
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  ::emu::g_cpasync_pending[::emu::emu_lin_tid()].push_back([dst, src, src_bytes] {
    if (src_bytes) memcpy(dst, src, 16); else memset(dst, 0, 16);
  });
}
static inline void cp_async_4(uint32_t dst_smem, const void* src) {
  if ((dst_smem & 3) || ((uintptr_t)src & 3)) ::emu::tc_fail("cp.async 4: dst %u / src %p must be 4-byte aligned", dst_smem, src);
  uint8_t* dst = ::emu::smem_ptr(dst_smem, 4);
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  ::emu::g_cpasync_pending[::emu::emu_lin_tid()].push_back([dst, src] { memcpy(dst, src, 4); });
}
// cp.async.wait_all: the calling thread's copies land now (the latest legal moment)
static inline void cp_async_wait_all() {
  std::vector<std::function<void()>> copies;
  {
    std::lock_guard<std::mutex> g(::emu::g_tc_mu);
    copies = std::move(::emu::g_cpasync_pending[::emu::emu_lin_tid()]);
    ::emu::g_cpasync_pending[::emu::emu_lin_tid()].clear();
  }
  for (auto& c : copies) c();
}
// flags between CTAs: thread blocks run one after another here, so these are plain atomics
static inline uint32_t ld_acquire_gpu(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void st_release_gpu(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
static inline void st_relaxed_gpu(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
static inline void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  auto& b = bar_at(bar);
  (void)b;
  auto copies = std::move(::emu::g_cpasync_pending[::emu::emu_lin_tid()]);
  ::emu::g_cpasync_pending[::emu::emu_lin_tid()].clear();
  ::emu::AsyncOp op;
  op.kind = 3;                 // cp.async group: runs its copies, then ARRIVES (no transaction bytes)
  op.bar = smem_u32(bar);
  op.tx = 0;
  op.run = [copies] { for (auto& c : copies) c(); };
  if (!::emu::g_async_late) {
    op.run();
    if (b.pending <= 0) ::emu::tc_fail("mbarrier at %u: more arrivals than its count %u", op.bar, b.expected);
    --b.pending;
    bar_check(b);
    return;
  }
  ::emu::g_copy_queue.push_back(std::move(op));
}

// ---- TMEM -----------------------------------------------------------------------------
static inline void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // one full warp
  __syncwarp();
  if ((::emu::emu_lin_tid() & 31) == 0) {
    if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) ::emu::tc_fail("tcgen05.alloc: %u columns", ncols);
    std::lock_guard<std::mutex> g(::emu::g_tc_mu);
    if (::emu::g_tmem_next + ncols > 512) ::emu::tc_fail("tcgen05.alloc: out of tensor memory");
    const uint32_t a = smem_u32(dst_smem);
    const uint32_t base = ::emu::g_tmem_next;  // lane 0, column base
    memcpy(::emu::smem_ptr(a, 4), &base, 4);
    ::emu::g_tmem_next += ncols;
    ::emu::g_tmem_live += ncols;
  }
  __syncwarp();
}
static inline void tmem_relinquish() {}
static inline void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  __syncwarp();
  if ((::emu::emu_lin_tid() & 31) == 0) {
    std::lock_guard<std::mutex> g(::emu::g_tc_mu);
    if ((taddr >> 16) != 0 || (taddr & 0xFFFF) + ncols > 512 || ncols > ::emu::g_tmem_live)
      ::emu::tc_fail("tcgen05.dealloc: bad range (address 0x%x, %u columns)", taddr, ncols);
    ::emu::g_tmem_live -= ncols;
  }
  __syncwarp();
}
static inline void fence_before_sync() {}
static inline void fence_after_sync() {}

// ---- descriptors ----------------------------------------------------------------------
static inline uint64_t desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1, kFmtTF32 = 2 };
static inline uint32_t idesc_f32acc(uint32_t ab_format, int M, int N) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct IDesc {
  int M, N;
  uint32_t afmt, bfmt;
};
static inline IDesc decode_idesc(uint32_t idesc, uint32_t kind_fmt_lo, uint32_t kind_fmt_hi) {
  IDesc d;
  d.N = (int)((idesc >> 17) & 0x3F) << 3;
  d.M = (int)((idesc >> 24) & 0x1F) << 4;
  d.afmt = (idesc >> 7) & 7;
  d.bfmt = (idesc >> 10) & 7;
  if (((idesc >> 4) & 3) != 1) ::emu::tc_fail("idesc: accumulator format is not f32");
  if ((idesc >> 15) & 3) ::emu::tc_fail("idesc: transposed (MN-major) operands are not modelled");
  if (d.afmt < kind_fmt_lo || d.afmt > kind_fmt_hi || d.bfmt < kind_fmt_lo || d.bfmt > kind_fmt_hi)
    ::emu::tc_fail("idesc: operand formats %u/%u do not belong to this .kind", d.afmt, d.bfmt);
  if (d.M != 128) ::emu::tc_fail("idesc: M = %d (only the M = 128 accumulator layout is modelled)", d.M);
  if (d.N < 16 || d.N > 256 || d.N % 16) ::emu::tc_fail("idesc: N = %d is not a valid shape for M = 128", d.N);
  return d;
}
// byte `b` of the 32-byte K slice of operand row r, K-major SWIZZLE_128B canonical layout
static inline const uint8_t* operand_byte(uint64_t desc, int r, int b) {
  if ((desc >> 61) != 2) ::emu::tc_fail("smem descriptor: swizzle mode %u (only SWIZZLE_128B is modelled)",
                                        (unsigned)(desc >> 61));
  if (((desc >> 46) & 3) != 1) ::emu::tc_fail("smem descriptor: version field must be 1 on sm_100");
  if ((desc >> 49) & 7) ::emu::tc_fail("smem descriptor: base-offset field set (operand base not 1024-B aligned?)");
  const uint32_t start = (uint32_t)(desc & 0x3FFF) << 4;
  const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3FFF) << 4;
  uint32_t lin = start + (uint32_t)(r >> 3) * sbo + (uint32_t)(r & 7) * 128u + (uint32_t)b;
  lin ^= ((lin >> 7) & 7u) << 4;  // the swizzle is a function of the absolute shared address bits
  return ::emu::smem_ptr(lin, 1);
}
static inline float tf32_operand(uint32_t bits) {  // the tensor core ignores the low 13 mantissa bits
  bits &= 0xFFFFE000u;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static inline float bf16_operand(uint16_t h, uint32_t fmt) {
  if (fmt == kFmtBF16) {
    uint32_t bits = (uint32_t)h << 16;
    float f;
    memcpy(&f, &bits, 4);
    return f;
  }
  // IEEE half
  const uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
  float v;
  if (e == 0) v = std::ldexp((float)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = std::ldexp((float)(m | 1024), (int)e - 25);
  return s ? -v : v;
}
static inline void check_d(uint32_t d_tmem, int N) {
  if ((d_tmem >> 16) != 0) ::emu::tc_fail("tcgen05.mma: accumulator address 0x%x has a lane offset", d_tmem);
  if ((d_tmem & 0xFFFF) + (uint32_t)N > ::emu::g_tmem_next)
    ::emu::tc_fail("tcgen05.mma: accumulator columns [%u, +%d) exceed the allocation (%u)", d_tmem & 0xFFFF, N,
                   ::emu::g_tmem_next);
}
// D[128 x N] (+)= A[128 x K] * B[N x K]^T, a(m, k) / b(n, k) supplied by the caller
template <typename FA, typename FB>
static inline void mma_core(uint32_t d_tmem, int N, int K, uint32_t accumulate, FA a, FB b) {
  check_d(d_tmem, N);
  const uint32_t c0 = d_tmem & 0xFFFF;
  std::vector<float> bv((size_t)N * K);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) bv[(size_t)n * K + k] = b(n, k);
  float av[16];
  for (int m = 0; m < 128; ++m) {
    for (int k = 0; k < K; ++k) av[k] = a(m, k);
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += av[k] * bv[(size_t)n * K + k];
      float d = 0.f;
      if (accumulate) memcpy(&d, &::emu::g_tmem[m][c0 + n], 4);
      d += acc;
      memcpy(&::emu::g_tmem[m][c0 + n], &d, 4);
    }
  }
  ::emu::g_mma_count.fetch_add(1);
}
static inline float smem_tf32(uint64_t desc, int r, int k) {
  uint32_t bits;
  const uint8_t* p = operand_byte(desc, r, 4 * k);
  memcpy(&bits, p, 4);  // 4-byte elements never straddle a 16-byte swizzle unit
  return tf32_operand(bits);
}
static inline float smem_h16(uint64_t desc, int r, int k, uint32_t fmt) {
  uint16_t h;
  memcpy(&h, operand_byte(desc, r, 2 * k), 2);
  return bf16_operand(h, fmt);
}

static inline void mma_issue(std::function<void()> fn) {
  if (!::emu::g_async_late) { fn(); return; }
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  ::emu::AsyncOp op;
  op.kind = 0;
  op.run = std::move(fn);
  ::emu::g_mma_queue.push_back(std::move(op));
}
static inline void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const IDesc I = decode_idesc(idesc, kFmtTF32, kFmtTF32);
  check_d(d_tmem, I.N);
  mma_issue([=] {
    mma_core(d_tmem, I.N, 8, accumulate, [&](int m, int k) { return smem_tf32(adesc, m, k); },
             [&](int n, int k) { return smem_tf32(bdesc, n, k); });
  });
}
static inline void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const IDesc I = decode_idesc(idesc, kFmtF16, kFmtBF16);
  check_d(d_tmem, I.N);
  mma_issue([=] {
    mma_core(d_tmem, I.N, 16, accumulate, [&](int m, int k) { return smem_h16(adesc, m, k, I.afmt); },
             [&](int n, int k) { return smem_h16(bdesc, n, k, I.bfmt); });
  });
}
static inline uint32_t desc_lo32(uint32_t smem_addr) { return (smem_addr & 0x3FFFFu) >> 4; }
static inline uint32_t desc_hi32_k_sw128() { return (uint32_t)(desc_k_sw128(0) >> 32); }
static inline void mma_f16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                              uint32_t accumulate) {
  // the low word carries the 14-bit start-address field (and the leading-dimension field above it, unused here):
  // an add that carried out of the address field would silently corrupt the descriptor on the device
  if ((a_lo >> 14) || (b_lo >> 14)) ::emu::tc_fail("mma_f16_lo: descriptor low word overflows the address field");
  mma_f16(d_tmem, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, accumulate);
}
static inline void check_a_tmem(uint32_t a_tmem, int cols) {
  if ((a_tmem >> 16) != 0) ::emu::tc_fail("tcgen05.mma: TMEM A address 0x%x has a lane offset", a_tmem);
  if ((a_tmem & 0xFFFF) + (uint32_t)cols > ::emu::g_tmem_next)
    ::emu::tc_fail("tcgen05.mma: TMEM A columns exceed the allocation");
}
static inline void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                               uint32_t accumulate) {
  const IDesc I = decode_idesc(idesc, kFmtTF32, kFmtTF32);
  check_a_tmem(a_tmem, 8);
  const uint32_t ac = a_tmem & 0xFFFF;
  check_d(d_tmem, I.N);
  mma_issue([=] {
    mma_core(d_tmem, I.N, 8, accumulate, [&](int m, int k) { return tf32_operand(::emu::g_tmem[m][ac + k]); },
             [&](int n, int k) { return smem_tf32(bdesc, n, k); });
  });
}
// A in tensor memory, 16-bit operands: two K elements per 32-bit column (low half = even k)
static inline void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                              uint32_t accumulate) {
  const IDesc I = decode_idesc(idesc, kFmtF16, kFmtBF16);
  check_a_tmem(a_tmem, 8);
  const uint32_t ac = a_tmem & 0xFFFF;
  check_d(d_tmem, I.N);
  mma_issue([=] {
    mma_core(d_tmem, I.N, 16, accumulate,
             [&](int m, int k) {
               const uint32_t w = ::emu::g_tmem[m][ac + (k >> 1)];
               return bf16_operand((uint16_t)((k & 1) ? (w >> 16) : (w & 0xFFFF)), I.afmt);
             },
             [&](int n, int k) { return smem_h16(bdesc, n, k, I.bfmt); });
  });
}
// the commit's arrival happens once every MMA issued before it has executed (see AsyncOp)
static inline void mma_commit(uint64_t* bar) {
  if (!::emu::g_async_late) { mbar_arrive(bar); return; }
  std::lock_guard<std::mutex> g(::emu::g_tc_mu);
  bar_at(bar);
  ::emu::AsyncOp op;
  op.kind = 1;
  op.bar = smem_u32(bar);
  ::emu::g_mma_queue.push_back(std::move(op));
}

static inline void tmem_lane_rule(uint32_t taddr, const char* what, uint32_t ncols = 16) {
  const int tid = ::emu::emu_lin_tid();
  const uint32_t lane0 = taddr >> 16;
  if (lane0 != (uint32_t)(((tid >> 5) & 3) * 32))
    ::emu::tc_fail("%s: warp %d may only touch TMEM lanes %d.., address names lane %u", what, tid >> 5,
                   ((tid >> 5) & 3) * 32, lane0);
  if ((taddr & 0xFFFF) + ncols > ::emu::g_tmem_next) ::emu::tc_fail("%s: columns exceed the allocation", what);
}
static inline void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  tmem_lane_rule(taddr, "tcgen05.ld");
  const int lane = (int)(taddr >> 16) + (::emu::emu_lin_tid() & 31);
  for (int i = 0; i < 16; ++i) v[i] = ::emu::g_tmem[lane][(taddr & 0xFFFF) + i];
}
static inline void tmem_ld_wait() {}
static inline void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  tmem_lane_rule(taddr, "tcgen05.st");
  const int lane = (int)(taddr >> 16) + (::emu::emu_lin_tid() & 31);
  for (int i = 0; i < 16; ++i) ::emu::g_tmem[lane][(taddr & 0xFFFF) + i] = v[i];
}
static inline void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  tmem_lane_rule(taddr, "tcgen05.st", 8);
  const int lane = (int)(taddr >> 16) + (::emu::emu_lin_tid() & 31);
  for (int i = 0; i < 8; ++i) ::emu::g_tmem[lane][(taddr & 0xFFFF) + i] = v[i];
}
static inline void tmem_st_wait() {}

static inline void pdl_launch_dependents() {}
static inline void pdl_wait() {}

static inline float4 ld_shared_v4(uint32_t addr) {
  if (addr & 15) ::emu::tc_fail("ld.shared.v4 at %u is not 16-byte aligned", addr);
  float4 v;
  memcpy(&v, ::emu::smem_ptr(addr, 16), 16);
  return v;
}
static inline void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  if (addr & 15) ::emu::tc_fail("st.shared.v4 at %u is not 16-byte aligned", addr);
  const float v[4] = {a, b, c, d};
  memcpy(::emu::smem_ptr(addr, 16), v, 16);
}
static inline void st_shared_v2_b32(uint32_t addr, uint32_t a, uint32_t b) {
  if (addr & 7) ::emu::tc_fail("st.shared.v2 at %u is not 8-byte aligned", addr);
  const uint32_t v[2] = {a, b};
  memcpy(::emu::smem_ptr(addr, 8), v, 8);
}
static inline void st_shared_v4_b32(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  if (addr & 15) ::emu::tc_fail("st.shared.v4 at %u is not 16-byte aligned", addr);
  const uint32_t v[4] = {a, b, c, d};
  memcpy(::emu::smem_ptr(addr, 16), v, 16);
}

static inline void named_bar_sync(int id, int nthreads) {
  unsigned gen;
  {
    std::lock_guard<std::mutex> g(::emu::g_tc_mu);
    auto& b = ::emu::g_named[id];
    gen = b.gen;
    if (++b.count == nthreads) {
      b.count = 0;
      ++b.gen;
      return;
    }
  }
  const auto t0 = std::chrono::steady_clock::now();
  for (long spin = 0;; ++spin) {
    {
      std::lock_guard<std::mutex> g(::emu::g_tc_mu);
      if (::emu::g_named[id].gen != gen) return;
    }
    ::emu::blocked_yield();
    if ((spin & 0xFFF) == 0xFFF && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30))
      ::emu::tc_fail("deadlock: named barrier %d (%d threads) not reached by everyone", id, nthreads);
  }
}

static inline float round_tf32(float x) {  // cvt.rna.tf32.f32: nearest, ties away from zero
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
// cvt.rn.bf16x2.f32: two floats -> packed bf16 pair (lo in the low half), nearest even
static inline uint16_t bf16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline uint32_t pack_bf16x2(float lo, float hi) {
  return (uint32_t)bf16_rn(lo) | ((uint32_t)bf16_rn(hi) << 16);
}
static inline float bf16_round(float x) {
  const uint32_t u = (uint32_t)bf16_rn(x) << 16;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

}  // namespace tc
}  // namespace msmd
