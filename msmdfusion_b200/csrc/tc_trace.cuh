// tc_trace.cuh -- per-role timeline of the warp-specialised tensor-core kernels, for finding which chain
// (gather latency, stage hand-off, weight bulk copy, MMA issue) sets the per-chunk period.  Compiled in ONLY
// with -DMSMD_TC_TRACE (`python -m msmdfusion_b200.build --trace` -> _C/libmsmd_b200_trace.so, used by
// tools/tc_trace.py); in the product library every macro below expands to nothing.
//
// Buffer (device, unsigned long long): kTrCtas records of kTrRecord words
//   head [16]: 0 clock at entry | 1 setup done | 2 main loop done (gather warp 0) | 4 epilogue done | 5 exit
//              6 %smid | 7 active chunks | 8 blockIdx.x | 9 %globaltimer at entry | 10 %globaltimer at exit
//   then [role][chunk it < kTrIts][phase < 4] = clock64():
//     role 0 / 1  gather warp 0 / 7, lane 0:  0 about to wait for the free stage | 1 stage free | 2 stored + arrived
//     role 2      weight loader:             0 about to wait for the free stage | 1 stage free, bulk copy issued
//     role 3      MMA issuer:                0 about to wait for the full stage | 1 stage full | 2 MMAs issued + commit
//   persistent kernel (spconv_fwd_sbp_kernel) only -- `it` = segment index:
//     role 4      epilogue warp 0:           0 about to wait for the accumulator | 1 accumulator complete |
//                                            2 hand-off flags seen (owner of a split tile) | 3 rows written
//     role 6      epilogue warp 0, it = 8 * segment + column step: 0 TMEM read done | 1 staged (after __syncwarp) |
//                                            2 staged block read back | 3 rows stored
//     role 5      segment loader:            0 about to wait for the free slot | 1 slot free | 2 pair rows landed | 3 published
// Traced CTAs: kTrCtas of them, evenly spaced over the grid.
#pragma once
#ifdef MSMD_TC_TRACE
namespace msmd {
constexpr int kTrCtas = 16, kTrRoles = 7, kTrIts = 128, kTrPhases = 4, kTrHead = 16;
constexpr int kTrRecord = kTrHead + kTrRoles * kTrIts * kTrPhases;
static __device__ unsigned long long* g_tc_trace = nullptr;
__device__ __forceinline__ unsigned long long* tc_trace_base() {
  unsigned long long* buf = g_tc_trace;
  if (!buf) return nullptr;
  const unsigned stride = gridDim.x > (unsigned)kTrCtas ? gridDim.x / kTrCtas : 1u;
  if (blockIdx.x % stride || blockIdx.x / stride >= (unsigned)kTrCtas) return nullptr;
  return buf + (size_t)(blockIdx.x / stride) * kTrRecord;
}
__device__ __forceinline__ unsigned long long tc_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned tc_smid() {
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}
static int tc_trace_set_impl(unsigned long long* buf) {
  return cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -2;
}
}  // namespace msmd
#define TC_TRACE_INIT() unsigned long long* const tr_base = ::msmd::tc_trace_base()
#define TC_TRACE_HEAD(slot, value) \
  do { if (tr_base) tr_base[slot] = (unsigned long long)(value); } while (0)
#define TC_TRACE(role, it, phase)                                                                        \
  do {                                                                                                   \
    if (tr_base && (role) >= 0 && (it) < ::msmd::kTrIts)                                                 \
      tr_base[::msmd::kTrHead + (((role) * ::msmd::kTrIts + (it)) * ::msmd::kTrPhases + (phase))] =      \
          (unsigned long long)clock64();                                                                 \
  } while (0)
#define TC_TRACE_ENTRY()                                         \
  do {                                                           \
    if (tr_base && threadIdx.x == 0) {                           \
      tr_base[0] = (unsigned long long)clock64();                \
      tr_base[6] = ::msmd::tc_smid();                            \
      tr_base[8] = blockIdx.x;                                   \
      tr_base[9] = ::msmd::tc_globaltimer();                     \
    }                                                            \
  } while (0)
#define TC_TRACE_EXIT()                                          \
  do {                                                           \
    if (tr_base && threadIdx.x == 0) {                           \
      tr_base[5] = (unsigned long long)clock64();                \
      tr_base[10] = ::msmd::tc_globaltimer();                    \
    }                                                            \
  } while (0)
#else
#define TC_TRACE_INIT() do {} while (0)
#define TC_TRACE_HEAD(slot, value) do {} while (0)
#define TC_TRACE(role, it, phase) do {} while (0)
#define TC_TRACE_ENTRY() do {} while (0)
#define TC_TRACE_EXIT() do {} while (0)
#endif
