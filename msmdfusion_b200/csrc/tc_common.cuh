// tc_common.cuh -- pieces shared by the tensor-core sparse-convolution kernels (spconv_tc.cu: tf32 x3;
// spconv_tc16.cu: bf16 / bf16 x3): tile constants, the active-chunk list, the TMEM epilogue.
#pragma once
#include "tc.cuh"

namespace msmd {

constexpr int kTcProducerWarps = 8;                       // gather warps (also the epilogue)
constexpr int kTcProducers = kTcProducerWarps * 32;       // 256 threads
constexpr int kTcThreads = kTcProducers + 64;             // + B-loader warp + MMA warp
constexpr int kTcM = 128;        // output voxels per tile (UMMA M)
constexpr int kTcKC = 32;        // floats per K chunk = one 128-byte swizzle row
constexpr int kTcABytes = kTcM * kTcKC * 4;  // 16 KB per A half (hi or lo)
constexpr int kTcDepth = 3;      // variant 2: K chunks whose gather loads are in flight per thread

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// Host-side tuning switches for A/B runs (msmd_spconv_tc_set_tuning; defaults = the measured heuristics):
//   [0] occupancy  0 auto | 1 one CTA per SM (whole shared memory, deeper pipeline) | 2 two CTAs per SM when they fit
//   [1] stage cap  2..4 (default 4)
//   [2] split-K    0 auto | 1 never | 2 whenever the kernel supports it
//   [3] chunk blocks per pipeline stage of the 16-bit kernel (variant 2): 0/1 one | 2 two (half the hand-offs)
//   [4] persistent split-operand kernel: weight of a tile's epilogue in K-chunk times, plus one (0 = by N)
extern int g_tc_tune[5];
// shared-memory budget of one CTA: `fits_half` = two pipeline stages fit in half of the SM
static inline int tc_smem_budget(bool fits_half, int tiles) {
  const int half = 112 * 1024, full = 224 * 1024;
  if (g_tc_tune[0] == 1) return full;
  if (g_tc_tune[0] == 2) return fits_half ? half : full;
  return (fits_half && tiles > kNumSMs) ? half : full;
}
static inline int tc_stage_cap() { return (g_tc_tune[1] >= 2 && g_tc_tune[1] <= 4) ? g_tc_tune[1] : 4; }

// Launch of a tensor-core conv kernel.  Product build: an ordinary triple-chevron launch.  -DMSMD_TC_PDL (debug build
// `python -m msmdfusion_b200.build --pdl`): programmatic stream serialization, so that consecutive layers of a
// chain overlap the next kernel's prologue (barrier init, TMEM allocation, pair-table load, active-chunk list)
// with this kernel's tail; the kernels order their data reads with tc::pdl_wait().
template <typename Kern, typename... Args>
static inline void tc_launch(Kern kern, int grid, int block, int smem, cudaStream_t stream, Args... args) {
#ifdef MSMD_TC_PDL
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, args...);
#else
  kern<<<grid, block, smem, stream>>>(args...);
#endif
}

// One lane polls the mbarrier, the warp follows through __syncwarp (31 fewer spinning lanes per
// warp: the producers are instruction-issue bound, profiles/r01e_ncu_full_spconv_tc_profileS.json).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) tc::mbar_wait(bar, parity);
  __syncwarp();
}
__device__ __forceinline__ void mbar_wait_warp_long(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) tc::mbar_wait_long(bar, parity);
  __syncwarp();
}

// Compact list of the K chunks this tile uses (a chunk is skipped when none of its kernel
// offsets has a pair in the tile).  Executed by warp 0; returns the count in *n_act_s.
__device__ __forceinline__ void tc_build_active_list(const int* used_s, int chunks, int cin_pad, int kvol,
                                                     int lane, unsigned short* alist, int* n_act_s,
                                                     int kc = kTcKC /* K elements per chunk */) {
  int cnt = 0;
  for (int base = 0; base < chunks; base += 32) {
    const int j = base + lane;
    int a = 0;
    if (j < chunks) {
      const int k_lo = (j * kc) / cin_pad;
      int k_hi = (j * kc + kc - 1) / cin_pad;
      if (k_hi > kvol - 1) k_hi = kvol - 1;
      for (int k = k_lo; k <= k_hi; ++k) a |= used_s[k];
    }
    const unsigned b = __ballot_sync(0xffffffffu, a != 0);
    if (a) alist[cnt + __popc(b & ((1u << lane) - 1u))] = (unsigned short)j;
    cnt += __popc(b);
  }
  if (lane == 0) *n_act_s = cnt;
}

// Epilogue shared by the kernel variants: executed by the 8 producer warps once the accumulator
// barrier fires.  Warp w may read TMEM lanes 32*(w&3)..+31 (= accumulator rows); the two
// warpgroups split the N columns.  y = relu(acc*scale + shift + residual), 16-byte stores.
__device__ __forceinline__ void tc_epilogue(int any_active, uint64_t* accum_bar, uint32_t tmem_base,
                                            int warp, int lane, int row0, int n_out, int cout, int N,
                                            const float* ss /* smem: scale[N] | shift[N] */,
                                            const float* __restrict__ residual, int relu,
                                            float* __restrict__ out, float* write_partial = nullptr,
                                            const float* add_partial = nullptr, int cat_cols = 0,
                                            const int* __restrict__ row_perm = nullptr) {
    if (any_active) {
      tc::mbar_wait(accum_bar, 0);
      tc::fence_after_sync();
    }
    const int quarter = warp & 3;             // TMEM lanes this warp may read: 32*quarter ..
    int o = row0 + quarter * 32 + lane;
    // mask-sorted tiles: the pair table was permuted so that rows with similar neighbour masks share a
    // tile; tile slot o holds output row row_perm[o] (the residual is read, the result written there)
    if (row_perm) o = (o < n_out) ? __ldg(row_perm + o) : n_out;
    const int nsteps = N / 16;
    const int step_lo = (warp >> 2) ? (nsteps + 1) / 2 : 0;  // two warpgroups split the columns
    const int step_hi = (warp >> 2) ? nsteps : (nsteps + 1) / 2;
    const bool vec_out = (cout % 4 == 0) && (((uintptr_t)out & 15) == 0) &&
                         (residual == nullptr || ((uintptr_t)residual & 15) == 0);
    for (int st = step_lo; st < step_hi; ++st) {
      const int c0 = st * 16;
      uint32_t acc[16];
      if (any_active) {
        tc::tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, acc);
        if (cat_cols) {  // concatenated-B mode: columns [N, 2N) hold the A_hi * B_lo term
          uint32_t acc2[16];
          tc::tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cat_cols + c0), acc2);
          tc::tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e)
            acc[e] = __float_as_uint(__uint_as_float(acc[e]) + __uint_as_float(acc2[e]));
        }
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0u;
      }
      const int prow = quarter * 32 + lane;  // row inside the tile (split-K hand-off buffer [128][N])
      if (write_partial) {  // first finisher of a split-K pair: raw partial sums, no epilogue
        float4* dst = (float4*)(write_partial + (size_t)prow * N + c0);
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          __stcg(dst + (e >> 2), make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]),
                                             __uint_as_float(acc[e + 2]), __uint_as_float(acc[e + 3])));
        continue;
      }
      if (add_partial) {  // second finisher: add the partner's partial sums (L2, never L1)
        const float4* src = (const float4*)(add_partial + (size_t)prow * N + c0);
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const float4 pv = __ldcg(src + (e >> 2));
          acc[e] = __float_as_uint(__uint_as_float(acc[e]) + pv.x);
          acc[e + 1] = __float_as_uint(__uint_as_float(acc[e + 1]) + pv.y);
          acc[e + 2] = __float_as_uint(__uint_as_float(acc[e + 2]) + pv.z);
          acc[e + 3] = __float_as_uint(__uint_as_float(acc[e + 3]) + pv.w);
        }
      }
      if (o < n_out) {
        float* orow = out + (size_t)o * cout;
        const float* rrow = residual ? residual + (size_t)o * cout : nullptr;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int co = c0 + e;
          if (co >= cout) break;
          float y[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            y[q] = __uint_as_float(acc[e + q]);
            y[q] = fmaf(y[q], ss[co + q], ss[N + co + q]);  // (1, 0) when no BatchNorm is folded in
          }
          if (vec_out) {
            if (rrow) {
              const float4 rv = __ldg((const float4*)(rrow + co));
              y[0] += rv.x; y[1] += rv.y; y[2] += rv.z; y[3] += rv.w;
            }
            if (relu) {
#pragma unroll
              for (int q = 0; q < 4; ++q) y[q] = fmaxf(y[q], 0.f);
            }
            *(float4*)(orow + co) = make_float4(y[0], y[1], y[2], y[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (co + q >= cout) break;
              float t = y[q];
              if (rrow) t += __ldg(rrow + co + q);
              if (relu) t = fmaxf(t, 0.f);
              orow[co + q] = t;
            }
          }
        }
      }
    }
}

}  // namespace msmd
