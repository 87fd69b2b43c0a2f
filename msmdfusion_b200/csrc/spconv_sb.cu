// spconv_sb.cu -- sparse convolution forward on tcgen05 with a SPLIT-BF16 OPERAND CACHE and an asynchronous gather.
//
// Arithmetic: exactly the bf16 x3 mode of spconv_tc16.cu -- x = hi + lo with hi = bf16(x), lo = bf16(x - hi);
// D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi on tcgen05.mma.kind::f16, fp32 accumulator in tensor memory, fused
// BatchNorm / residual / ReLU epilogue.  What changes is WHERE the split happens.
//
// r02a measurements (profiles/r02a_*): every earlier kernel pays ~1.2 us per 32-element gather step whatever
// Cout, MMA kind, occupancy or stage count is -- the gather warps' load -> convert -> st.shared -> fence -> arrive
// chain, not the tensor pipe and not HBM.  A row of activations is written ONCE by its producing layer and
// gathered 3..27 times by the next one, so the hi / lo split belongs to the producer:
//
//   * every activation tensor the convolutions read exists (also) as a "split image"
//         xs[row] = [ hi(0..C8) | lo(0..C8) ]   bf16,  C8 = round_up(C, 8),  4*C8 bytes per row (= the fp32 row),
//     written by the epilogue of the layer that produced it next to the fp32 tensor (which stays the API-visible
//     result and the residual operand), or by msmd_split_bf16 for a network input;
//   * the gather is then a pure copy: cp.async (LDGSTS, 16 bytes, L2 -> shared memory, no registers) straight into
//     the K-major SWIZZLE_128B operand tile, missing pairs zero-filled by src-size 0; each of the 256 producer
//     threads ties its copies to the stage's "full" mbarrier with cp.async.mbarrier.arrive.noinc and runs ahead as
//     far as free stages allow -- up to `stages` x 32 KB of gather traffic in flight per CTA and no thread ever
//     waits for its own loads;
//   * weights: the [hi | lo] 128-byte-swizzled images of spconv_tc16.cu (one bulk copy per chunk), with the input
//     channels padded to 8 so that a 16-byte piece never straddles two kernel offsets.
//
// K chunk = 64 bf16 elements (one swizzle row); K index = k * C8 + c.  Roles as in the other kernels: warps 0-7
// producers + epilogue, warp 8 weight loader, warp 9 TMEM owner + MMA issuer.  For 2N <= 256 the adjacent
// B_hi | B_lo images are one 2N-row operand (2 MMAs per k-step, A_hi read from shared memory once).
//
// Reference semantics: spconv v2.1.21 SparseConvolution.forward (API copy bug_fix/conv.py:382-447); parity is
// checked against the oracle and the reference's vendored spconv-1.x goldens (tests/test_gpu_parity.py).
#include "tc_common.cuh"
#include "tc_trace.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace msmd {

constexpr int kSbKC = 64;                 // bf16 K elements per chunk
constexpr int kSbABytes = kTcM * 128;     // 16 KB per A image (hi or lo)

struct SbGeom {
  int cin_pad, N, chunks;
};
static bool sb_geom(int cout, int kvol, int cin, SbGeom& g) {
  if (cout < 1 || cout > 256 || kvol < 1 || kvol > 32 || cin < 1 || cin > 4096) return false;
  g.cin_pad = round_up(cin, 8);
  g.N = round_up(cout, 16);
  g.chunks = (kvol * g.cin_pad + kSbKC - 1) / kSbKC;
  return true;
}

struct SbLayout {
  int stage_bytes, stages, pair_off, act_off, bar_off, total;
};
static SbLayout sb_layout(int N, int kvol, int chunks, int tiles) {
  SbLayout L;
  L.stage_bytes = 2 * kSbABytes + 2 * N * 128;   // A hi | A lo | B hi | B lo
  const int misc = round_up(kvol * kTcM * 4, 16) + round_up(2 * chunks, 16) + 256 + 8 * N;
  const int budget = tc_smem_budget(2 * L.stage_bytes + misc + 1024 <= 112 * 1024, tiles);
  L.stages = (budget - misc - 1024) / L.stage_bytes;
  if (L.stages > 6) L.stages = 6;
  if (g_tc_tune[1] >= 2 && g_tc_tune[1] <= 6 && L.stages > g_tc_tune[1]) L.stages = g_tc_tune[1];
  L.pair_off = L.stages * L.stage_bytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(2 * chunks, 16);
  L.total = L.bar_off + 256 + 8 * N + 1024;
  return L;
}

// Epilogue: TMEM -> registers -> y = relu(acc*scale + shift + residual) -> fp32 rows and / or the split image.
__device__ __forceinline__ void sb_epilogue(int any_active, uint64_t* accum_bar, uint32_t tmem_base, int warp,
                                            int lane, int row0, int n_out, int cout, int cout_pad, int N,
                                            const float* ss, const float* __restrict__ residual, int relu,
                                            float* __restrict__ out, uint16_t* __restrict__ out_s, int cat_cols) {
  if (any_active) {
    tc::mbar_wait(accum_bar, 0);
    tc::fence_after_sync();
  }
  const int quarter = warp & 3;
  const int o = row0 + quarter * 32 + lane;
  const int nsteps = N / 16;
  const int step_lo = (warp >> 2) ? (nsteps + 1) / 2 : 0;
  const int step_hi = (warp >> 2) ? nsteps : (nsteps + 1) / 2;
  const bool vec_out = (cout % 4 == 0) && (out == nullptr || ((uintptr_t)out & 15) == 0) &&
                       (residual == nullptr || ((uintptr_t)residual & 15) == 0);
  for (int st = step_lo; st < step_hi; ++st) {
    const int c0 = st * 16;
    uint32_t acc[16];
    if (any_active) {
      tc::tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, acc);
      if (cat_cols) {
        uint32_t acc2[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cat_cols + c0), acc2);
        tc::tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = __float_as_uint(__uint_as_float(acc[e]) + __uint_as_float(acc2[e]));
      }
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0u;
    }
    if (o >= n_out) continue;
    float y[16];
    const float* rrow = residual ? residual + (size_t)o * cout : nullptr;
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
      const int co = c0 + e;
#pragma unroll
      for (int q = 0; q < 4; ++q) y[e + q] = fmaf(__uint_as_float(acc[e + q]), ss[co + q], ss[N + co + q]);
      if (rrow && co < cout) {
        if (vec_out) {
          const float4 rv = __ldg((const float4*)(rrow + co));
          y[e] += rv.x; y[e + 1] += rv.y; y[e + 2] += rv.z; y[e + 3] += rv.w;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (co + q < cout) y[e + q] += __ldg(rrow + co + q);
        }
      }
      if (relu) {
#pragma unroll
        for (int q = 0; q < 4; ++q) y[e + q] = fmaxf(y[e + q], 0.f);
      }
    }
    if (out) {
      float* orow = out + (size_t)o * cout;
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const int co = c0 + e;
        if (co >= cout) break;
        if (vec_out) {
          *(float4*)(orow + co) = make_float4(y[e], y[e + 1], y[e + 2], y[e + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (co + q < cout) orow[co + q] = y[e + q];
        }
      }
    }
    if (out_s) {
      // split image of the row: hi at [0, cout_pad), lo at [cout_pad, 2*cout_pad); channels in [cout, cout_pad) are zero
      uint16_t* srow = out_s + (size_t)o * (2 * cout_pad);
#pragma unroll
      for (int e = 0; e < 16; e += 8) {
        const int co = c0 + e;
        if (co >= cout_pad) break;
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = (co + 2 * q < cout) ? y[e + 2 * q] : 0.f;
          const float b = (co + 2 * q + 1 < cout) ? y[e + 2 * q + 1] : 0.f;
          h[q] = tc::pack_bf16x2(a, b);
          l[q] = tc::pack_bf16x2(a - __uint_as_float(h[q] << 16), b - __uint_as_float(h[q] & 0xFFFF0000u));
        }
        *(uint4*)(srow + co) = make_uint4(h[0], h[1], h[2], h[3]);
        *(uint4*)(srow + cout_pad + co) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
  }
}

__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_sb_kernel(const uint16_t* __restrict__ xs, const uint16_t* __restrict__ wpk,
                     const int* __restrict__ pair, int n_out, int cin_pad, uint32_t cin_magic, int cout, int cout_pad,
                     int N, int kvol,
                     int chunks, int stages, int stage_bytes, int pair_off, int act_off, int bar_off, int tmem_cols,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ residual, int relu, float* __restrict__ out,
                     uint16_t* __restrict__ out_s, int cat) {
  extern __shared__ uint8_t smem_raw[];
  // round up to 1024 bytes INSIDE the shared window (an integer round-trip of the generic address makes ptxas
  // emit generic LD.E instead of LDS for everything derived from it: seen in the r02f SASS of the producer loop)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  int* pair_s = (int*)(smem + pair_off);
  unsigned short* alist = (unsigned short*)(smem + act_off);
  uint64_t* full_bar = (uint64_t*)(smem + bar_off);
  uint64_t* empty_bar = full_bar + 6;
  uint64_t* accum_bar = full_bar + 12;
  uint32_t* tmem_ptr_s = (uint32_t*)(full_bar + 13);
  int* n_act_s = (int*)(full_bar + 13) + 1;
  int* used_s = (int*)(full_bar + 14);  // [kvol <= 32]
  float* ss = (float*)(smem + bar_off + 256);
  for (int c = threadIdx.x; c < N; c += kTcThreads) {
    ss[c] = (scale && c < cout) ? __ldg(scale + c) : 1.f;
    ss[N + c] = (shift && c < cout) ? __ldg(shift + c) : 0.f;
  }

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcM;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  tc::pdl_launch_dependents();

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], kTcProducers + 1);  // one cp.async-completion arrival per producer thread + the B copy
      tc::mbar_init(&empty_bar[s], 1);                // one tcgen05.commit
    }
    tc::mbar_init(accum_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == kTcProducerWarps + 1) {
    tc::tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
    tc::tmem_relinquish();
  }
  for (int k = warp; k < kvol; k += kTcThreads / 32) {
    bool any = false;
#pragma unroll
    for (int q = 0; q < kTcM / 32; ++q) {
      const int r = lane + 32 * q;
      const int o = row0 + r;
      const int p = (o < n_out) ? __ldg(pair + (size_t)k * n_out + o) : -1;
      pair_s[k * kTcM + r] = p;
      any |= p >= 0;
    }
    const unsigned b = __ballot_sync(0xffffffffu, any);
    if (lane == 0) used_s[k] = b != 0;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 0) tc_build_active_list(used_s, chunks, cin_pad, kvol, lane, alist, n_act_s, kSbKC);
  __syncthreads();
  const int n_act = *n_act_s;
  const int any_active = n_act > 0;
  const uint32_t tmem_base = *tmem_ptr_s;
  if (tid == 0) { TC_TRACE_HEAD(1, clock64()); TC_TRACE_HEAD(7, n_act); }
  tc::pdl_wait();

  if (warp < kTcProducerWarps) {
    // ===== A producers: 16-byte pieces, cp.async straight into the swizzled tile ===================
    // thread (q, rbase): piece q (8 K elements) of rows rbase + 32*i of every chunk
    const int q = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    const size_t row_elems = (size_t)2 * cin_pad;
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);  // traced gather warps
    (void)tr_role;
    int s = 0;
    uint32_t ph = 1u;   // "empty" barriers: the first pass through the ring finds every stage free
    // The loop body is one latency chain per warp (chunk id -> kernel offset -> 4 pair indices -> 4 addresses ->
    // 8 copies), and only two producer warps share a scheduler: keep it short and wide -- the division by cin_pad is
    // a multiply-high by a host-computed reciprocal, the four index loads are independent, invalid pairs are
    // zero-filled by predication (src-size 0), no branches.  (An attempt to skip slots that stay empty, r02f, cost
    // more in divergence than the zero-fill copies it saved.)
    for (int t = 0; t < n_act; ++t) {
      if (lane == 0) TC_TRACE(tr_role, t, 0);
      mbar_wait_warp(&empty_bar[s], ph, lane);
      if (lane == 0) TC_TRACE(tr_role, t, 1);
      const uint32_t kk0 = (uint32_t)alist[t] * kSbKC + (uint32_t)q * 8u;
      const uint32_t k = __umulhi(kk0, cin_magic);          // kk0 / cin_pad (exact for kk0 < 2^16, checked on the host)
      const uint32_t c0 = kk0 - k * (uint32_t)cin_pad;
      const bool kvalid = k < (uint32_t)kvol;
      const int* prow = pair_s + (kvalid ? k : 0u) * kTcM + rbase;
      int idx[kTcM / 32];
#pragma unroll
      for (int i = 0; i < kTcM / 32; ++i) idx[i] = prow[32 * i];
      const uint32_t a_hi = tc::smem_u32(smem) + (uint32_t)s * (uint32_t)stage_bytes + (uint32_t)(rbase * 128) +
                            (uint32_t)((q ^ (rbase & 7)) << 4);   // rows rbase + 32 i share (row & 7)
#pragma unroll
      for (int i = 0; i < kTcM / 32; ++i) {
        const bool v = kvalid && idx[i] >= 0;
        const uint16_t* src = xs + (v ? (size_t)idx[i] * row_elems + c0 : 0);
        const uint32_t nb = v ? 16u : 0u;
        tc::cp_async_16(a_hi + (uint32_t)(i * 32 * 128), src, nb);
        tc::cp_async_16(a_hi + (uint32_t)(i * 32 * 128) + kSbABytes, v ? src + cin_pad : src, nb);
      }
      tc::cp_async_mbar_arrive_noinc(&full_bar[s]);
      if (lane == 0) TC_TRACE(tr_role, t, 2);
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    if (tid == 0) TC_TRACE_HEAD(2, clock64());
    sb_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, cout_pad, N, ss, residual, relu,
                out, out_s, cat ? N : 0);
    if (tid == 0) TC_TRACE_HEAD(4, clock64());
  } else if (warp == kTcProducerWarps) {
    // ===== B loader ================================================================================
    {
      const uint32_t bytes = (uint32_t)(2 * N * 128);
      int s = 0;
      uint32_t ph = 1u;
      for (int t = 0; t < n_act; ++t) {
        if (lane == 0) TC_TRACE(2, t, 0);
        tc::mbar_wait(&empty_bar[s], ph);
        if (lane == 0) TC_TRACE(2, t, 1);
        const uint8_t* src = (const uint8_t*)wpk + (size_t)alist[t] * bytes;
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
          tc::bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kSbABytes, src, bytes, &full_bar[s]);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===== MMA issuer ==============================================================================
    // One lane issues everything, so its instruction count is a per-chunk floor: descriptors are 32-bit low
    // words over one constant high word, advanced by integer adds (tc::mma_f16_lo).
    // The whole warp runs the loop (converged); the asynchronous-unit instructions are issued by the elected lane.
    {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtBF16, kTcM, N);
      const uint32_t idesc2 = tc::idesc_f32acc(tc::kFmtBF16, kTcM, 2 * N);
      const uint32_t dhi = tc::desc_hi32_k_sw128();
      const uint32_t lo0 = tc::desc_lo32(tc::smem_u32(smem));
      const uint32_t stage_lo = (uint32_t)stage_bytes >> 4, a_lo_off = (uint32_t)kSbABytes >> 4;
      const uint32_t b_hi_off = (uint32_t)(2 * kSbABytes) >> 4, b_lo_off = b_hi_off + (((uint32_t)N * 128u) >> 4);
      uint32_t accumulate = 0;
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < n_act; ++t) {
        if (lane == 0) TC_TRACE(3, t, 0);
        tc::mbar_wait(&full_bar[s], ph);
        if (lane == 0) TC_TRACE(3, t, 1);
        tc::fence_proxy_async();   // cp.async wrote the A tile through the generic proxy
        tc::fence_after_sync();
        const uint32_t a = lo0 + (uint32_t)s * stage_lo;
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kSbKC / 16; ++ks) {
            const uint32_t ah = a + (uint32_t)ks * 2u;   // 16 bf16 = 32 bytes along K = 2 descriptor units
            if (cat) {
              tc::mma_f16_lo(tmem_base, ah, ah + b_hi_off, dhi, idesc2, accumulate);   // D[:, 0:2N] += A_hi * [B_hi; B_lo]
              tc::mma_f16_lo(tmem_base, ah + a_lo_off, ah + b_hi_off, dhi, idesc, 1u);  // D[:, 0:N]  += A_lo * B_hi
            } else {
              tc::mma_f16_lo(tmem_base, ah + a_lo_off, ah + b_hi_off, dhi, idesc, accumulate);
              tc::mma_f16_lo(tmem_base, ah, ah + b_lo_off, dhi, idesc, 1u);
              tc::mma_f16_lo(tmem_base, ah, ah + b_hi_off, dhi, idesc, 1u);
            }
            accumulate = 1u;
          }
          tc::mma_commit(&empty_bar[s]);
          if (t == n_act - 1) tc::mma_commit(accum_bar);   // accumulator complete -> epilogue
        }
        __syncwarp();
        accumulate = 1u;
        if (lane == 0) TC_TRACE(3, t, 2);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  TC_TRACE_EXIT();
  if (warp == kTcProducerWarps + 1) tc::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ======================================================================================================
// Persistent, work-balanced variant (r02l).
//
// The kernel above is at the measured tensor rate INSIDE its main loop (0.53 us per 128 x 384 x 64 chunk of the
// 128 -> 128 layers = 11 TFLOP/s per SM = MEASURED_PEAKS' dense bf16 rate / 148), and loses the rest around it
// (profiles/r02j ncu list): a grid of 169 tiles on 148 SMs runs two waves for 1.14 waves of work, and the
// 1.5 us set-up + 8 us epilogue of every tile are serial with its 24 us main loop.  This variant keeps the tile
// code and changes the schedule:
//   * one CTA per SM slot, each given an equal, contiguous range of the launch's units -- per tile `epi_units` for
//     its epilogue plus one per K chunk that has a pair in the tile (tile masks; all chunks without them).  A tile
//     that straddles a range boundary is finished by the CTA holding its FIRST chunk (its owner, which reaches it
//     last); the CTAs holding the rest reach their fragment first, and hand the owner raw fp32 partial sums
//     through an L2-resident slot + one release/acquire flag per epilogue warp.  The owner adds the partials in
//     CTA order, so the result does not depend on timing;
//   * two accumulators in tensor memory and dedicated epilogue warps: the epilogue of segment i runs under the
//     main loop of segment i + 1; a loader warp fetches the next segment's pair rows (cp.async) and builds its
//     active-chunk list one segment ahead.
// Warp roles: 0-7 A producers, 8 B loader, 9 MMA issuer + TMEM owner, 10 segment loader, 11.. epilogue (4 or 8).
// ======================================================================================================
constexpr int kSbpWarpB = kTcProducerWarps;
constexpr int kSbpWarpMma = kTcProducerWarps + 1;
constexpr int kSbpWarpSeg = kTcProducerWarps + 2;
constexpr int kSbpWarpEpi = kTcProducerWarps + 3;
constexpr int kSbpBarBytes = 320;   // 20 mbarriers | tmem pointer | segment records [2][8] | start record [4]

struct SbpLayout {
  int stage_bytes, stages, pair_off, pair_bytes, act_off, act_bytes, kmask_off, bar_off, total;
};
static SbpLayout sbp_layout(int N, int kvol, int chunks, int budget, int epi_warps) {
  SbpLayout L;
  L.stage_bytes = 2 * kSbABytes + 2 * N * 128;
  L.pair_bytes = round_up(kvol * kTcM * 4, 16);
  if (L.pair_bytes < epi_warps * 2048) L.pair_bytes = epi_warps * 2048;   // doubles as the epilogue's staging blocks
  L.act_bytes = round_up(2 * chunks, 16);
  const int kmask_bytes = round_up(4 * chunks, 16);
  const int misc = 2 * L.pair_bytes + 2 * L.act_bytes + kmask_bytes + kSbpBarBytes + 8 * N + 1024;
  L.stages = (budget - misc) / L.stage_bytes;
  if (L.stages > 6) L.stages = 6;
  if (g_tc_tune[1] >= 2 && g_tc_tune[1] <= 6 && L.stages > g_tc_tune[1]) L.stages = g_tc_tune[1];
  if (L.stages < 0) L.stages = 0;
  L.pair_off = L.stages * L.stage_bytes;
  L.act_off = L.pair_off + 2 * L.pair_bytes;
  L.kmask_off = L.act_off + 2 * L.act_bytes;
  L.bar_off = L.kmask_off + kmask_bytes;
  L.total = L.bar_off + kSbpBarBytes + 8 * N + 1024;
  return L;
}

__device__ __forceinline__ void flag_wait(const uint32_t* flag) {
  if (tc::ld_acquire_gpu(flag) != 0u) return;
  const long long t0 = clock64();
  while (tc::ld_acquire_gpu(flag) == 0u) {
    if (clock64() - t0 > 4000000000LL) __trap();   // a protocol bug must surface as an error, never hang the GPU
  }
}

template <int kEpiWarps>   // 8: one CTA per SM | 4: two CTAs per SM (<= 68 registers)
__global__ void __launch_bounds__((kSbpWarpEpi + kEpiWarps) * 32, kEpiWarps == 4 ? 2 : 1)
spconv_fwd_sbp_kernel(const uint16_t* __restrict__ xs, const uint16_t* __restrict__ wpk,
                      const int* __restrict__ pair, const int* __restrict__ row_perm,
                      const uint32_t* __restrict__ tile_mask, int n_out, int tiles, int cin_pad,
                      uint32_t cin_magic, int cout, int cout_pad, int N, int kvol, int chunks, int stages,
                      int stage_bytes, int pair_off, int pair_bytes, int act_off, int act_bytes, int kmask_off,
                      int bar_off, int tmem_cols, const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ residual, int relu, float* __restrict__ out,
                      uint16_t* __restrict__ out_s, int cat, int epi_units, float* __restrict__ ws,
                      uint32_t* __restrict__ flags) {
  constexpr int epi_warps = kEpiWarps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = (uint64_t*)(smem + bar_off);
  uint64_t* empty_bar = full_bar + 6;
  uint64_t* acc_full = full_bar + 12;    // [2] MMA issuer -> epilogue: accumulator complete
  uint64_t* acc_empty = full_bar + 14;   // [2] epilogue -> MMA issuer: accumulator read
  uint64_t* seg_full = full_bar + 16;    // [2] segment loader -> everyone: pair rows + active list + record ready
  uint64_t* seg_empty = full_bar + 18;   // [2] everyone -> segment loader
  uint32_t* tmem_ptr_s = (uint32_t*)(full_bar + 20);
  // segment record: 0 chunks to multiply | 1 tile | 2 bit 0 = holds the tile's first unit (owner), bit 1 = reaches its
  // last unit | 3 unit index one past the tile's last unit | 4 units of the CTA's range this segment covers
  int* seg_info = (int*)(smem + bar_off + 176);   // [2][8]
  int* start_s = (int*)(smem + bar_off + 240);    // total units | first tile of the range | first unit inside it
  uint32_t* kmask_s = (uint32_t*)(smem + kmask_off);   // [chunks]: the kernel offsets chunk j covers, one bit each
  float* ss = (float*)(smem + bar_off + kSbpBarBytes);
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    ss[c] = (scale && c < cout) ? __ldg(scale + c) : 1.f;
    ss[N + c] = (shift && c < cout) ? __ldg(shift + c) : 0.f;
  }
  for (int j = threadIdx.x; j < chunks; j += blockDim.x) {
    const int k_lo = (j * kSbKC) / cin_pad;
    int k_hi = (j * kSbKC + kSbKC - 1) / cin_pad;
    if (k_hi > kvol - 1) k_hi = kvol - 1;
    kmask_s[j] = (0xffffffffu >> (31 - k_hi)) & (0xffffffffu << k_lo);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  // Programmatic dependent launch: the next kernel of the stream may start its CTAs as SMs free up; everything up
  // to griddep_wait() below reads only what no kernel of the chain writes (weights' scale / shift, the rulebook, its
  // tile masks), so this kernel's own prologue overlaps its predecessor's tail the same way.
  tc::griddep_launch_dependents();

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], kTcProducers + 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], (uint32_t)epi_warps);
      tc::mbar_init(&seg_full[b], 1);
      tc::mbar_init(&seg_empty[b], (uint32_t)(kTcProducerWarps + 2 + epi_warps));
    }
    tc::fence_mbar_init();
  }
  if (warp == kSbpWarpMma) {
    tc::tmem_alloc(tmem_ptr_s, (uint32_t)(2 * tmem_cols));
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr_s;

  // ---- the CTA's share of the launch --------------------------------------------------------------------------
  // A tile is worth `epi_units` (its epilogue, in chunk times; they come first and belong to the tile's owner) plus
  // one unit per K chunk it has to multiply: all of them without `tile_mask`, otherwise the chunks in which at least
  // one of its rows has a pair.  Every CTA takes units [W*cta/G, W*(cta+1)/G) of the launch's W.  With tile masks the
  // position of the first unit needs the prefix over the tiles: block-wide, once, in the (still idle) pipeline stages.
  const int whole = (cin_pad % kSbKC == 0) ? cin_pad / kSbKC : 0;   // chunks per kernel offset when they nest
  auto tile_units = [&](int t) {
    const uint32_t m = __ldg(tile_mask + t);
    int cnt = 0;
    if (whole) cnt = __popc(m) * whole;
    else for (int j = 0; j < chunks; ++j) cnt += (kmask_s[j] & m) != 0u;
    return epi_units + cnt;
  };
  int total_units, u_begin, u_end;
  if (tile_mask) {
    int* part = (int*)smem;
    const int T = (int)blockDim.x;
    const int per = (tiles + T - 1) / T;
    const int tb = min(tiles, tid * per), te = min(tiles, tb + per);
    int sum = 0;
    for (int t = tb; t < te; ++t) sum += tile_units(t);
    part[tid] = sum;
    __syncthreads();
    if (warp == 0) {
      const int per_l = (T + 31) / 32;
      const int ib = min(T, lane * per_l), ie = min(T, ib + per_l);
      int acc = 0;
      for (int i = ib; i < ie; ++i) { const int v = part[i]; part[i] = acc; acc += v; }
      int incl = acc;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const int excl = incl - acc;
      for (int i = ib; i < ie; ++i) part[i] += excl;
      if (lane == 31) start_s[0] = incl;
    }
    __syncthreads();
    total_units = start_s[0];
    u_begin = (int)((long long)total_units * cta / G);
    u_end = (int)((long long)total_units * (cta + 1) / G);
    const int ex = part[tid];
    if (u_begin < u_end && u_begin >= ex && u_begin < ex + sum) {
      int t = tb, acc = ex;
      for (;;) {
        const int w = tile_units(t);
        if (acc + w > u_begin) break;
        acc += w;
        ++t;
      }
      start_s[1] = t;
      start_s[2] = u_begin - acc;
    }
    __syncthreads();
  } else {
    total_units = tiles * (epi_units + chunks);
    u_begin = (int)((long long)total_units * cta / G);
    u_end = (int)((long long)total_units * (cta + 1) / G);
  }
  if (tid == 0) { TC_TRACE_HEAD(1, clock64()); TC_TRACE_HEAD(7, u_end - u_begin); }
  // activations, residual, outputs and the hand-off slots are touched only behind this point; the segment loader
  // (pair rows, tile masks: geometry-stream data the launching stream already waited for) runs ahead
  if (warp != kSbpWarpSeg) tc::griddep_wait();

// every role walks the CTA's segments in step: the loader publishes a record per segment, the others read it
#define SBP_SEGMENT_LOOP for (int u = u_begin, seg = 0; u < u_end; ++seg)
#define SBP_SEGMENT_RECORD(wait_stmt)                             \
  const int b = seg & 1;                                          \
  const uint32_t sph = (uint32_t)(seg >> 1) & 1u;                 \
  wait_stmt;                                                      \
  const int* const si = seg_info + b * 8;                         \
  const int n_act = si[0], tile = si[1], sflags = si[2], tile_end = si[3]; \
  u += si[4];                                                     \
  (void)tile; (void)sflags; (void)tile_end; (void)sph

  if (warp < kTcProducerWarps) {
    // ===== A producers (the loop body of spconv_fwd_sb_kernel; the stage ring runs on across segments) =====
    const int q = tid & 7;
    const int rbase = tid >> 3;
    const size_t row_elems = (size_t)2 * cin_pad;
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);
    (void)tr_role;
    int s = 0, it = 0;
    uint32_t ph = 1u;
    SBP_SEGMENT_LOOP {
      SBP_SEGMENT_RECORD(mbar_wait_warp_long(&seg_full[b], sph, lane));
      const int* pair_s = (const int*)(smem + pair_off + b * pair_bytes);
      const unsigned short* alist = (const unsigned short*)(smem + act_off + b * act_bytes);
      for (int t = 0; t < n_act; ++t, ++it) {
        if (lane == 0) TC_TRACE(tr_role, it, 0);
        mbar_wait_warp(&empty_bar[s], ph, lane);
        if (lane == 0) TC_TRACE(tr_role, it, 1);
        const uint32_t kk0 = (uint32_t)alist[t] * kSbKC + (uint32_t)q * 8u;
        const uint32_t k = __umulhi(kk0, cin_magic);
        const uint32_t c0 = kk0 - k * (uint32_t)cin_pad;
        const bool kvalid = k < (uint32_t)kvol;
        const int* prow = pair_s + (kvalid ? k : 0u) * kTcM + rbase;
        int idx[kTcM / 32];
#pragma unroll
        for (int i = 0; i < kTcM / 32; ++i) idx[i] = prow[32 * i];
        const uint32_t a_hi = tc::smem_u32(smem) + (uint32_t)s * (uint32_t)stage_bytes + (uint32_t)(rbase * 128) +
                              (uint32_t)((q ^ (rbase & 7)) << 4);
#pragma unroll
        for (int i = 0; i < kTcM / 32; ++i) {
          const bool v = kvalid && idx[i] >= 0;
          const uint16_t* src = xs + (v ? (size_t)idx[i] * row_elems + c0 : 0);
          const uint32_t nb = v ? 16u : 0u;
          tc::cp_async_16(a_hi + (uint32_t)(i * 32 * 128), src, nb);
          tc::cp_async_16(a_hi + (uint32_t)(i * 32 * 128) + kSbABytes, v ? src + cin_pad : src, nb);
        }
        tc::cp_async_mbar_arrive_noinc(&full_bar[s]);
        if (lane == 0) TC_TRACE(tr_role, it, 2);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&seg_empty[b]);
    }
    if (tid == 0) TC_TRACE_HEAD(2, clock64());
  } else if (warp == kSbpWarpB) {
    // ===== B loader ================================================================================
    const uint32_t bytes = (uint32_t)(2 * N * 128);
    int s = 0, it = 0;
    uint32_t ph = 1u;
    SBP_SEGMENT_LOOP {
      SBP_SEGMENT_RECORD(mbar_wait_warp_long(&seg_full[b], sph, lane));
      const unsigned short* alist = (const unsigned short*)(smem + act_off + b * act_bytes);
      for (int t = 0; t < n_act; ++t, ++it) {
        if (lane == 0) TC_TRACE(2, it, 0);
        tc::mbar_wait(&empty_bar[s], ph);
        if (lane == 0) TC_TRACE(2, it, 1);
        const uint8_t* src = (const uint8_t*)wpk + (size_t)alist[t] * bytes;
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
          tc::bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kSbABytes, src, bytes, &full_bar[s]);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&seg_empty[b]);
    }
  } else if (warp == kSbpWarpMma) {
    // ===== MMA issuer ==============================================================================
    const uint32_t idesc = tc::idesc_f32acc(tc::kFmtBF16, kTcM, N);
    const uint32_t idesc2 = tc::idesc_f32acc(tc::kFmtBF16, kTcM, 2 * N);
    const uint32_t dhi = tc::desc_hi32_k_sw128();
    const uint32_t lo0 = tc::desc_lo32(tc::smem_u32(smem));
    const uint32_t stage_lo = (uint32_t)stage_bytes >> 4, a_lo_off = (uint32_t)kSbABytes >> 4;
    const uint32_t b_hi_off = (uint32_t)(2 * kSbABytes) >> 4, b_lo_off = b_hi_off + (((uint32_t)N * 128u) >> 4);
    int s = 0, it = 0;
    uint32_t ph = 0;
    SBP_SEGMENT_LOOP {
      SBP_SEGMENT_RECORD(tc::mbar_wait_long(&seg_full[b], sph));
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&seg_empty[b]);
      tc::mbar_wait_long(&acc_empty[b], sph ^ 1u);   // the epilogue has read the accumulator used two segments ago
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + (uint32_t)(b * tmem_cols);
      if (n_act == 0) {
        if (lane == 0) tc::mbar_arrive(&acc_full[b]);   // nothing to multiply: the epilogue takes zeros
        __syncwarp();
      }
      uint32_t accumulate = 0;
      for (int t = 0; t < n_act; ++t, ++it) {
        if (lane == 0) TC_TRACE(3, it, 0);
        tc::mbar_wait(&full_bar[s], ph);
        if (lane == 0) TC_TRACE(3, it, 1);
        tc::fence_proxy_async();
        tc::fence_after_sync();
        const uint32_t a = lo0 + (uint32_t)s * stage_lo;
        if (tc::elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kSbKC / 16; ++ks) {
            const uint32_t ah = a + (uint32_t)ks * 2u;
            if (cat) {
              tc::mma_f16_lo(d_tmem, ah, ah + b_hi_off, dhi, idesc2, accumulate);
              tc::mma_f16_lo(d_tmem, ah + a_lo_off, ah + b_hi_off, dhi, idesc, 1u);
            } else {
              tc::mma_f16_lo(d_tmem, ah + a_lo_off, ah + b_hi_off, dhi, idesc, accumulate);
              tc::mma_f16_lo(d_tmem, ah, ah + b_lo_off, dhi, idesc, 1u);
              tc::mma_f16_lo(d_tmem, ah, ah + b_hi_off, dhi, idesc, 1u);
            }
            accumulate = 1u;
          }
          tc::mma_commit(&empty_bar[s]);
          if (t == n_act - 1) tc::mma_commit(&acc_full[b]);
        }
        __syncwarp();
        accumulate = 1u;
        if (lane == 0) TC_TRACE(3, it, 2);
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kSbpWarpSeg) {
    // ===== segment loader: pair rows of the tile, the list of its K chunks to multiply, the segment record =====
    int t = 0, off = 0;
    if (tile_mask) {
      t = start_s[1];
      off = start_s[2];
    } else if (u_begin < u_end) {
      t = u_begin / (epi_units + chunks);
      off = u_begin - t * (epi_units + chunks);
    }
    SBP_SEGMENT_LOOP {
      const int b = seg & 1;
      if (lane == 0) TC_TRACE(5, seg, 0);
      mbar_wait_warp_long(&seg_empty[b], ((uint32_t)(seg >> 1) & 1u) ^ 1u, lane);
      if (lane == 0) TC_TRACE(5, seg, 1);
      int* pair_s = (int*)(smem + pair_off + b * pair_bytes);
      unsigned short* alist = (unsigned short*)(smem + act_off + b * act_bytes);
      const int row0 = t * kTcM;
      for (int k = 0; k < kvol; ++k) {
#pragma unroll
        for (int i = 0; i < kTcM / 32; ++i) {
          const int r = lane + 32 * i;
          const int o = row0 + r;
          if (o < n_out) tc::cp_async_4(tc::smem_u32(pair_s + k * kTcM + r), pair + (size_t)k * n_out + o);
          else pair_s[k * kTcM + r] = -1;
        }
      }
      uint32_t m;
      if (tile_mask) {
        m = __ldg(tile_mask + t);   // the list does not wait for the pair rows
      } else {
        tc::cp_async_wait_all();
        __syncwarp();
        m = 0u;
        for (int k = 0; k < kvol; ++k) {
          bool any = false;
#pragma unroll
          for (int i = 0; i < kTcM / 32; ++i) any |= pair_s[k * kTcM + lane + 32 * i] >= 0;
          m |= (any ? 1u : 0u) << k;
        }
        m = __reduce_or_sync(0xffffffffu, m);
      }
      // the tile's units: [0, epi_units) its epilogue, then one per chunk of its list; the segment takes
      // [off, off + take) of them, i.e. list positions [lo, hi)
      int units_t, take, n_list = 0;
      if (tile_mask) {
        int total = 0;
        for (int base = 0; base < chunks; base += 32) {
          const int j = base + lane;
          total += __popc(__ballot_sync(0xffffffffu, j < chunks && (kmask_s[j] & m) != 0u));
        }
        units_t = epi_units + total;
        take = min(units_t - off, u_end - u);
        const int lo = max(off, epi_units) - epi_units, hi = off + take - epi_units;
        int ord = 0;   // position in the tile's list of chunks with a pair
        for (int base = 0; base < chunks; base += 32) {
          const int j = base + lane;
          const bool a = j < chunks && (kmask_s[j] & m) != 0u;
          const unsigned bal = __ballot_sync(0xffffffffu, a);
          const int mine = ord + __popc(bal & ((1u << lane) - 1u));
          if (a && mine >= lo && mine < hi) alist[mine - lo] = (unsigned short)j;
          ord += __popc(bal);
        }
        n_list = hi > lo ? hi - lo : 0;
      } else {
        units_t = epi_units + chunks;
        take = min(units_t - off, u_end - u);
        const int lo = max(off, epi_units) - epi_units, hi = off + take - epi_units;
        for (int base = lo; base < hi; base += 32) {
          const int j = base + lane;
          const bool a = j < hi && (kmask_s[j] & m) != 0u;
          const unsigned bal = __ballot_sync(0xffffffffu, a);
          if (a) alist[n_list + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)j;
          n_list += __popc(bal);
        }
      }
      if (tile_mask) {
        tc::cp_async_wait_all();
        __syncwarp();
      }
      if (lane == 0) TC_TRACE(5, seg, 2);
      if (lane == 0) {
        int* si = seg_info + b * 8;
        si[0] = n_list;
        si[1] = t;
        si[2] = (off == 0 ? 1 : 0) | (off + take == units_t ? 2 : 0);
        si[3] = u - off + units_t;
        si[4] = take;
      }
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(&seg_full[b]); TC_TRACE(5, seg, 3); }
      u += take;
      if (off + take == units_t) { ++t; off = 0; } else { off += take; }
    }
  } else if (warp < kSbpWarpEpi + epi_warps) {
    // ===== epilogue ================================================================================
    const int ew = warp - kSbpWarpEpi;
    const int quarter = warp & 3;                 // the TMEM lanes (accumulator rows) this warp may read
    const int halves = epi_warps >> 2;            // 1 or 2 warps per quarter split the columns
    const int half = ew >> 2;
    const int nsteps = N / 16;
    const int per = (nsteps + halves - 1) / halves;
    const int step_lo = half * per, step_hi = min(nsteps, step_lo + per);
    const bool vec_out = (cout % 4 == 0) && (out == nullptr || ((uintptr_t)out & 15) == 0) &&
                         (residual == nullptr || ((uintptr_t)residual & 15) == 0);
    const int prow = quarter * 32 + lane;
    const size_t slot = (size_t)kTcM * N;         // floats per partial slot, [16-column step][row][16]
    const int total_w = total_units;
    SBP_SEGMENT_LOOP {
      SBP_SEGMENT_RECORD(tc::mbar_wait_long(&seg_full[b], sph));
      if (ew == 0 && lane == 0) TC_TRACE(4, seg, 0);
      tc::mbar_wait_long(&acc_full[b], sph);
      tc::fence_after_sync();
      if (ew == 0 && lane == 0) TC_TRACE(4, seg, 1);
      const uint32_t t_addr = tmem_base + (uint32_t)(b * tmem_cols) + ((uint32_t)(quarter * 32) << 16);
      const bool owner = (sflags & 1) != 0;
      const bool split = owner && (sflags & 2) == 0;   // other CTAs hold the rest of this tile
      int slot_o = tile * kTcM + prow;
      int o = slot_o;
      if (row_perm) o = (slot_o < n_out) ? __ldg(row_perm + slot_o) : n_out;
      // the row-contiguous half of the epilogue: this lane's four rows, and the staging block -- the segment's pair
      // rows are dead once its accumulator is complete, so the block lives in that buffer (held until the end of
      // this epilogue: the segment loader refills it only after this warp's seg_empty arrival)
      int orow[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = tile * kTcM + quarter * 32 + (lane >> 2) + 8 * i;
        orow[i] = sl < n_out ? (row_perm ? __ldg(row_perm + sl) : sl) : n_out;
      }
      const uint32_t stg = tc::smem_u32(smem + pair_off + b * pair_bytes) + (uint32_t)(ew * 2048);
      // the CTAs after this one whose (non-empty) range starts inside the tile hold its other fragments
#define SBP_FOR_CONTRIBUTORS                                                                         \
  for (int cc = cta + 1, cs = 0;                                                                     \
       cc < G && (cs = (int)((long long)total_w * cc / G)) < tile_end; ++cc)                        \
    if ((int)((long long)total_w * (cc + 1) / G) > cs)
      if (split) {
        // wait (once per contributor) before the column loop; the flags are consumed below, after the reads
        SBP_FOR_CONTRIBUTORS {
          if (lane == 0) flag_wait(flags + (size_t)cc * 8 + ew);
        }
        __syncwarp();
      }
      if (ew == 0 && lane == 0) TC_TRACE(4, seg, 2);
      for (int st = step_lo; st < step_hi; ++st) {
        const int c0 = st * 16;
        uint32_t acc[16];
        if (n_act > 0) {
          tc::tmem_ld16(t_addr + (uint32_t)c0, acc);
          if (cat) {
            uint32_t acc2[16];
            tc::tmem_ld16(t_addr + (uint32_t)(N + c0), acc2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = __float_as_uint(__uint_as_float(acc[e]) + __uint_as_float(acc2[e]));
          }
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[e] = 0u;
        }
        if (ew == 0 && lane == 0) TC_TRACE(6, seg * 8 + (st - step_lo), 0);
        if (!owner) {   // raw partial sums for the tile's owner
          float4* dst = (float4*)(ws + (size_t)cta * slot + ((size_t)st * kTcM + prow) * 16);
#pragma unroll
          for (int e = 0; e < 16; e += 4)
            __stcg(dst + (e >> 2), make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]),
                                               __uint_as_float(acc[e + 2]), __uint_as_float(acc[e + 3])));
          continue;
        }
        if (split) {
          SBP_FOR_CONTRIBUTORS {
            const float4* src = (const float4*)(ws + (size_t)cc * slot + ((size_t)st * kTcM + prow) * 16);
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 pv = __ldcg(src + (e >> 2));
              acc[e] = __float_as_uint(__uint_as_float(acc[e]) + pv.x);
              acc[e + 1] = __float_as_uint(__uint_as_float(acc[e + 1]) + pv.y);
              acc[e + 2] = __float_as_uint(__uint_as_float(acc[e + 2]) + pv.z);
              acc[e + 3] = __float_as_uint(__uint_as_float(acc[e + 3]) + pv.w);
            }
          }
        }
        if (vec_out) {
          // lane = row  ->  lane = (row, 16-byte piece) through this warp's 2 KB staging block, so that every global
          // access of the epilogue is row-contiguous: 8 rows x 64 bytes per instruction instead of 32 rows x 16 bytes
          // (the r02n timeline: 8-9.5 us of epilogue per 128 x 128 tile, all of it LSU wavefronts of strided accesses)
#pragma unroll
          for (int pc = 0; pc < 4; ++pc)
            tc::st_shared_v4(stg + (uint32_t)(lane * 64) + (uint32_t)((pc ^ ((lane >> 1) & 3)) << 4),
                             __uint_as_float(acc[4 * pc]), __uint_as_float(acc[4 * pc + 1]),
                             __uint_as_float(acc[4 * pc + 2]), __uint_as_float(acc[4 * pc + 3]));
          __syncwarp();
          if (ew == 0 && lane == 0) TC_TRACE(6, seg * 8 + (st - step_lo), 1);
          const int pc = lane & 3;
          const int co = c0 + 4 * pc;
          if (co < cout) {
            const float4 sc = *(const float4*)(ss + co), sh = *(const float4*)(ss + N + co);
            float4 rv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
              rv[i] = (residual && orow[i] < n_out) ? __ldg((const float4*)(residual + (size_t)orow[i] * cout + co))
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = (lane >> 2) + 8 * i;
              const float4 v = tc::ld_shared_v4(stg + (uint32_t)(r * 64) + (uint32_t)((pc ^ ((r >> 1) & 3)) << 4));
              if (orow[i] >= n_out) continue;
              float y0 = fmaf(v.x, sc.x, sh.x) + rv[i].x, y1 = fmaf(v.y, sc.y, sh.y) + rv[i].y;
              float y2 = fmaf(v.z, sc.z, sh.z) + rv[i].z, y3 = fmaf(v.w, sc.w, sh.w) + rv[i].w;
              if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f); }
              if (out) *(float4*)(out + (size_t)orow[i] * cout + co) = make_float4(y0, y1, y2, y3);
              if (out_s) {
                uint16_t* srow = out_s + (size_t)orow[i] * (2 * cout_pad) + co;
                const uint32_t h01 = tc::pack_bf16x2(y0, y1), h23 = tc::pack_bf16x2(y2, y3);
                const uint32_t l01 = tc::pack_bf16x2(y0 - __uint_as_float(h01 << 16), y1 - __uint_as_float(h01 & 0xFFFF0000u));
                const uint32_t l23 = tc::pack_bf16x2(y2 - __uint_as_float(h23 << 16), y3 - __uint_as_float(h23 & 0xFFFF0000u));
                *(uint2*)srow = make_uint2(h01, h23);
                *(uint2*)(srow + cout_pad) = make_uint2(l01, l23);
                if (co + 4 == cout && cout_pad > cout) {   // channel padding of the split image stays zero
                  *(uint2*)(srow + 4) = make_uint2(0u, 0u);
                  *(uint2*)(srow + cout_pad + 4) = make_uint2(0u, 0u);
                }
              }
            }
          }
          __syncwarp();   // the block is rewritten by the next step
          if (ew == 0 && lane == 0) TC_TRACE(6, seg * 8 + (st - step_lo), 3);
          continue;
        }
        if (o >= n_out) continue;
        float y[16];
        const float* rrow = residual ? residual + (size_t)o * cout : nullptr;
#pragma unroll
        for (int e = 0; e < 16; e += 4) {
          const int co = c0 + e;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) y[e + qq] = fmaf(__uint_as_float(acc[e + qq]), ss[co + qq], ss[N + co + qq]);
          if (rrow && co < cout) {
            if (vec_out) {
              const float4 rv = __ldg((const float4*)(rrow + co));
              y[e] += rv.x; y[e + 1] += rv.y; y[e + 2] += rv.z; y[e + 3] += rv.w;
            } else {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq)
                if (co + qq < cout) y[e + qq] += __ldg(rrow + co + qq);
            }
          }
          if (relu) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) y[e + qq] = fmaxf(y[e + qq], 0.f);
          }
        }
        if (out) {
          float* orow = out + (size_t)o * cout;
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const int co = c0 + e;
            if (co >= cout) break;
            if (vec_out) {
              *(float4*)(orow + co) = make_float4(y[e], y[e + 1], y[e + 2], y[e + 3]);
            } else {
#pragma unroll
              for (int qq = 0; qq < 4; ++qq)
                if (co + qq < cout) orow[co + qq] = y[e + qq];
            }
          }
        }
        if (out_s) {
          uint16_t* srow = out_s + (size_t)o * (2 * cout_pad);
#pragma unroll
          for (int e = 0; e < 16; e += 8) {
            const int co = c0 + e;
            if (co >= cout_pad) break;
            uint32_t h[4], l[4];
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              const float av = (co + 2 * qq < cout) ? y[e + 2 * qq] : 0.f;
              const float bv = (co + 2 * qq + 1 < cout) ? y[e + 2 * qq + 1] : 0.f;
              h[qq] = tc::pack_bf16x2(av, bv);
              l[qq] = tc::pack_bf16x2(av - __uint_as_float(h[qq] << 16), bv - __uint_as_float(h[qq] & 0xFFFF0000u));
            }
            *(uint4*)(srow + co) = make_uint4(h[0], h[1], h[2], h[3]);
            *(uint4*)(srow + cout_pad + co) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
      }
      // the accumulator is free again; publish / consume the hand-off flags of this warp's share
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&acc_empty[b]);
        tc::mbar_arrive(&seg_empty[b]);
      }
      if (ew == 0 && lane == 0) TC_TRACE(4, seg, 3);
      if (!owner) {
        __threadfence();
        __syncwarp();
        if (lane == 0) tc::st_release_gpu(flags + (size_t)cta * 8 + ew, 1u);
      } else if (split) {
        __syncwarp();
        if (lane == 0) {
          SBP_FOR_CONTRIBUTORS {
            tc::st_relaxed_gpu(flags + (size_t)cc * 8 + ew, 0u);   // back to zero for the next launch on this stream
          }
        }
      }
#undef SBP_FOR_CONTRIBUTORS
    }
  }
#undef SBP_SEGMENT_LOOP
#undef SBP_SEGMENT_RECORD

  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) TC_TRACE_HEAD(4, clock64());
  TC_TRACE_EXIT();
  if (warp == kSbpWarpMma) tc::tmem_dealloc(tmem_base, (uint32_t)(2 * tmem_cols));
}

// fp32 rows -> split image [hi | lo] with the channel count padded to 8 (zeros)
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ x, long long n, int c, int c_pad, uint16_t* __restrict__ xs) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, pair of channels)
  const int half = c_pad / 2;
  if (t >= n * half) return;
  const long long r = t / half;
  const int ch = (int)(t - r * half) * 2;
  const float a = ch < c ? x[r * c + ch] : 0.f;
  const float b = ch + 1 < c ? x[r * c + ch + 1] : 0.f;
  const uint32_t h = tc::pack_bf16x2(a, b);
  const uint32_t l = tc::pack_bf16x2(a - __uint_as_float(h << 16), b - __uint_as_float(h & 0xFFFF0000u));
  *(uint32_t*)(xs + r * (2 * c_pad) + ch) = h;
  *(uint32_t*)(xs + r * (2 * c_pad) + c_pad + ch) = l;
}

// Packed weight image: [chunk j][image: hi, lo][n < N][64 bf16, 16-byte units swizzled by (n & 7)], K = k*cin_pad + c
__global__ void __launch_bounds__(256)
sb_pack_weight_kernel(const float* __restrict__ w, int cout, int kvol, int cin, int cin_pad, int N, int chunks,
                      uint16_t* __restrict__ packed) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)chunks * N * kSbKC;
  if (t >= total) return;
  const int kk = (int)(t % kSbKC);
  const int n = (int)((t / kSbKC) % N);
  const int j = (int)(t / ((size_t)kSbKC * N));
  const int K = j * kSbKC + kk;
  const int k = K / cin_pad, c = K - k * cin_pad;
  float val = 0.f;
  if (k < kvol && c < cin && n < cout) val = w[((size_t)n * kvol + k) * cin + c];
  const uint32_t hi = tc::pack_bf16x2(val, 0.f) & 0xFFFFu;
  const float lo = val - __uint_as_float(hi << 16);
  const size_t blk = (size_t)N * kSbKC;
  const size_t off = (size_t)n * kSbKC + (size_t)((((kk >> 3) ^ (n & 7)) << 3) + (kk & 7));
  packed[((size_t)j * 2 + 0) * blk + off] = (uint16_t)hi;
  packed[((size_t)j * 2 + 1) * blk + off] = (uint16_t)(tc::pack_bf16x2(lo, 0.f) & 0xFFFFu);
}

}  // namespace msmd

using namespace msmd;

#ifdef MSMD_TC_TRACE
extern "C" MSMD_API int msmd_sb_trace_set(unsigned long long* buf) { return tc_trace_set_impl(buf); }
#endif

// ---- persistent variant: schedule switch + the per-stream hand-off workspace -------------------------------
// 0 = default (persistent, work shares from the rulebooks' tile masks) | 1 = one tile per CTA (spconv_fwd_sb_kernel) |
// 2 = persistent with nominal shares (callers that keep rulebooks do not build tile masks)
static int g_sb_variant = 0;
extern "C" MSMD_API int msmd_spconv_sb_set_variant(int variant) {
  MSMD_REQUIRE(variant >= 0 && variant <= 2, "spconv_sb_set_variant: 0 (default), 1 (tile per CTA) or 2 (persistent, no masks)");
  g_sb_variant = variant;
  return MSMD_OK;
}
extern "C" MSMD_API int msmd_spconv_sb_uses_tile_masks(void) { return g_sb_variant == 0 ? 1 : 0; }
// Programmatic dependent launch of the persistent kernel (A/B switch; default OFF).  r02r on a B200: back-to-back
// launches of one layer get 6 % faster with it (the prologue hides under the predecessor's tail), but the whole L / LC
// steps get 14 % / 3 % SLOWER: with no gap left between the conv kernels, the rulebook kernels of the geometry stream
// (which later layers wait for) find no free SM -- a 608-thread CTA leaves room for nothing else.
static int g_sb_pdl = 0;
extern "C" MSMD_API int msmd_spconv_sb_set_pdl(int enable) {
  g_sb_pdl = enable ? 1 : 0;
  return MSMD_OK;
}

namespace {
// Launches on one stream are serial, so one slot set per (device, stream) is enough; the flags are zero between
// launches (their consumer resets them).  Allocated on first use, grown when a launch needs more.
struct SbpWorkspace {
  float* ws = nullptr;
  size_t ws_bytes = 0;
  uint32_t* flags = nullptr;
  int flag_ctas = 0;
};
std::mutex g_sbp_mu;
std::map<std::pair<int, cudaStream_t>, SbpWorkspace> g_sbp_ws;

int sbp_workspace(cudaStream_t stream, int ctas, size_t ws_bytes, SbpWorkspace* out) {
  int dev = 0;
  MSMD_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_sbp_mu);
  SbpWorkspace& w = g_sbp_ws[std::make_pair(dev, stream)];
  if (w.ws_bytes < ws_bytes) {
    if (w.ws) {
      MSMD_CUDA_OK(cudaStreamSynchronize(stream));   // earlier launches on the stream may still use the old slots
      MSMD_CUDA_OK(cudaFree(w.ws));
      w.ws = nullptr;
      w.ws_bytes = 0;
    }
    MSMD_CUDA_OK(cudaMalloc((void**)&w.ws, ws_bytes));
    w.ws_bytes = ws_bytes;
  }
  if (w.flag_ctas < ctas) {
    if (w.flags) {
      MSMD_CUDA_OK(cudaStreamSynchronize(stream));
      MSMD_CUDA_OK(cudaFree(w.flags));
      w.flags = nullptr;
      w.flag_ctas = 0;
    }
    const int n = ctas < 2 * kNumSMs ? 2 * kNumSMs : ctas;
    MSMD_CUDA_OK(cudaMalloc((void**)&w.flags, (size_t)n * 8 * sizeof(uint32_t)));
    MSMD_CUDA_OK(cudaMemsetAsync(w.flags, 0, (size_t)n * 8 * sizeof(uint32_t), stream));   // ordered before the launch
    w.flag_ctas = n;
  }
  *out = w;
  return MSMD_OK;
}
}  // namespace

extern "C" MSMD_API int msmd_split_width(int channels) { return 2 * round_up(channels > 0 ? channels : 1, 8); }

extern "C" MSMD_API int msmd_split_bf16(const float* x, int n, int channels, void* xs, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(n >= 0 && channels >= 1, "split_bf16: bad sizes");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(x && xs, "split_bf16: null pointer");
  MSMD_REQUIRE(((uintptr_t)xs & 15) == 0, "split_bf16: the split image must be 16-byte aligned");
  const int c_pad = round_up(channels, 8);
  const long long total = (long long)n * (c_pad / 2);
  split_bf16_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(x, (long long)n, channels, c_pad, (uint16_t*)xs);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API size_t msmd_spconv_sb_packed_bytes(int cout, int kvol, int cin) {
  SbGeom g;
  if (!sb_geom(cout, kvol, cin, g)) return 0;
  return (size_t)g.chunks * 2 * g.N * kSbKC * sizeof(uint16_t);
}

extern "C" MSMD_API int msmd_spconv_sb_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                                   void* packed, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SbGeom g;
  MSMD_REQUIRE(sb_geom(cout, kvol, cin, g), "spconv_sb: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(weight_krsc && packed, "spconv_sb_pack_weight: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed & 15) == 0, "spconv_sb_pack_weight: packed must be 16-byte aligned");
  const size_t total = (size_t)g.chunks * g.N * kSbKC;
  sb_pack_weight_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(weight_krsc, cout, kvol, cin, g.cin_pad,
                                                                             g.N, g.chunks, (uint16_t*)packed);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

// Which schedule a shape gets under the default variant (r02q per-layer A/B on the LC scene, profiles/): the persistent
// kernel wins wherever a tile's main loop is long against its fixed costs -- N >= 64 and >= 16 K chunks per tile
// (Cin >= 38 at 27 offsets); the narrow and the 3-offset layers keep one tile per CTA, where two CTAs share an SM.
static bool sb_use_persistent(const SbGeom& g) {
  if (g_sb_variant == 1) return false;
  if (g_sb_variant == 2) return true;
  return g.N >= 64 && g.chunks >= 16;
}

// Launch of the persistent kernel: occupancy (1 or 2 CTAs per SM), grid = the SM slots (never more CTAs than
// units of work), equal unit ranges.
static int sbp_launch(const SbGeom& g, const void* features_split, const void* packed_sb, const int* pair_fwd,
                      const int* row_perm, const uint32_t* tile_mask, int n_out, int cout, int kvol, const float* scale,
                      const float* shift,
                      const float* residual, int relu, float* out, void* out_split, cudaStream_t stream) {
  const int tiles = ceil_div(n_out, kTcM);
  const long long units = (long long)tiles * (g.chunks + 8);
  MSMD_REQUIRE(units < (1ll << 31), "spconv_fwd_sb: too many (tile, chunk) units");
  const int cat = (2 * g.N <= 256) ? 1 : 0;
  int tmem_cols = 32;
  while (tmem_cols < (cat ? 2 * g.N : g.N)) tmem_cols <<= 1;
  // two CTAs per SM when each still gets >= 3 stages in half of the shared memory and half of tensor memory, and
  // there is more than one SM's worth of work per slot; the tuning switch [0] forces either
  const SbpLayout half = sbp_layout(g.N, kvol, g.chunks, 112 * 1024, 4);
  bool two = half.stages >= 3 && 2 * tmem_cols <= 256 && units >= 4ll * 2 * kNumSMs;
  if (g_tc_tune[0] == 1) two = false;
  if (g_tc_tune[0] == 2) two = half.stages >= 2 && 2 * tmem_cols <= 256;
  const SbpLayout L = two ? half : sbp_layout(g.N, kvol, g.chunks, 227 * 1024, 8);
  MSMD_REQUIRE(L.stages >= 2, "spconv_fwd_sb: tile does not fit in shared memory");
  MSMD_REQUIRE(tiles <= L.stages * L.stage_bytes / 4, "spconv_fwd_sb: too many tiles for the in-kernel prefix");
  const int epi_warps = two ? 4 : 8;
  // what a tile's epilogue costs in chunk times (r02o timelines: ~4 us against 0.55 us per chunk at N = 128, ~1.5 us
  // against 0.45 us at N = 32); the tuning switch [4] overrides it for A/B runs
  int epi_units = g.N / 16;
  if (epi_units < 2) epi_units = 2;
  if (epi_units > 8) epi_units = 8;
  if (g_tc_tune[4] >= 1 && g_tc_tune[4] <= 64) epi_units = g_tc_tune[4] - 1;
  const int threads = (kSbpWarpEpi + epi_warps) * 32;
  static bool attr_set = false;
  static int resident[2] = {0, 0};   // CTAs per SM the hardware really co-schedules, [one, two]
  if (!attr_set) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(spconv_fwd_sbp_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MSMD_CUDA_OK(cudaFuncSetAttribute(spconv_fwd_sbp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set = true;
  }
  if (resident[two] == 0) {
    int n = 0;
    if (two) MSMD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spconv_fwd_sbp_kernel<4>, threads, (size_t)L.total));
    else MSMD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, spconv_fwd_sbp_kernel<8>, threads, (size_t)L.total));
    MSMD_REQUIRE(n >= 1, "spconv_fwd_sb: the persistent kernel does not fit on an SM");
    resident[two] = n > 2 ? 2 : n;
  }
  long long slots = (long long)kNumSMs * (two ? resident[1] : 1);
  // a range shorter than a few chunks is all hand-off: never more CTAs than units / 4
  if (slots > (units + 3) / 4) slots = (units + 3) / 4;
  if (slots < 1) slots = 1;
  const int grid = (int)slots;
  SbpWorkspace w;
  const int rc = sbp_workspace(stream, grid, (size_t)grid * kTcM * g.N * sizeof(float), &w);
  if (rc != MSMD_OK) return rc;
  const uint32_t cin_magic = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)g.cin_pad - 1) / (uint64_t)g.cin_pad);
  for (uint32_t kk = 0; kk < (uint32_t)g.chunks * kSbKC; kk += 8) {
    MSMD_REQUIRE((uint32_t)(((uint64_t)kk * cin_magic) >> 32) == kk / (uint32_t)g.cin_pad,
                 "spconv_fwd_sb: reciprocal of cin_pad %d is inexact at %u", g.cin_pad, kk);
  }
#define MSMD_SBP_ARGS                                                                                             \
  (const uint16_t*)features_split, (const uint16_t*)packed_sb, pair_fwd, row_perm, tile_mask, n_out, tiles,         \
      g.cin_pad, cin_magic, cout, round_up(cout, 8), g.N, kvol, g.chunks, L.stages, L.stage_bytes, L.pair_off,       \
      L.pair_bytes, L.act_off, L.act_bytes, L.kmask_off, L.bar_off, tmem_cols, scale, shift, residual, relu, out,    \
      (uint16_t*)out_split, cat, epi_units, w.ws, w.flags
#ifdef MSMD_EMUL
  if (two) spconv_fwd_sbp_kernel<4><<<grid, threads, L.total, stream>>>(MSMD_SBP_ARGS);
  else spconv_fwd_sbp_kernel<8><<<grid, threads, L.total, stream>>>(MSMD_SBP_ARGS);
#else
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)L.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_sb_pdl ? 1 : 0;
    const cudaError_t launch_err = two ? cudaLaunchKernelEx(&cfg, spconv_fwd_sbp_kernel<4>, MSMD_SBP_ARGS)
                                       : cudaLaunchKernelEx(&cfg, spconv_fwd_sbp_kernel<8>, MSMD_SBP_ARGS);
    MSMD_CUDA_OK(launch_err);
  }
#endif
#undef MSMD_SBP_ARGS
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_fwd_sb(const void* features_split, int n_in, const void* packed_sb,
                                           const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                           const float* scale, const float* shift, const float* residual, int relu,
                                           float* out, void* out_split, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  SbGeom g;
  MSMD_REQUIRE(sb_geom(cout, kvol, cin, g), "spconv_fwd_sb: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0, "spconv_fwd_sb: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd_sb: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(n_in > 0, "spconv_fwd_sb: output rows without input rows");
  MSMD_REQUIRE(features_split && packed_sb && pair_fwd && (out || out_split), "spconv_fwd_sb: null pointer");
  MSMD_REQUIRE(((uintptr_t)features_split & 15) == 0 && ((uintptr_t)packed_sb & 15) == 0 &&
                   ((uintptr_t)out_split & 15) == 0,
               "spconv_fwd_sb: split images and packed weights must be 16-byte aligned");
  const int tiles = ceil_div(n_out, kTcM);
  if (sb_use_persistent(g))
    return sbp_launch(g, features_split, packed_sb, pair_fwd, nullptr, nullptr, n_out, cout, kvol, scale, shift,
                      residual, relu, out, out_split, stream);
  const SbLayout L = sb_layout(g.N, kvol, g.chunks, tiles);
  MSMD_REQUIRE(L.stages >= 2, "spconv_fwd_sb: tile does not fit in shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(spconv_fwd_sb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // kk0 / cin_pad as __umulhi(kk0, magic): exact for every K index of the packed image (< 2^16), verified here once
  // per shape class -- the loop is over at most chunks * 8 values and runs on the host
  const uint32_t cin_magic = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)g.cin_pad - 1) / (uint64_t)g.cin_pad);
  for (uint32_t kk = 0; kk < (uint32_t)g.chunks * kSbKC; kk += 8) {
    MSMD_REQUIRE((uint32_t)(((uint64_t)kk * cin_magic) >> 32) == kk / (uint32_t)g.cin_pad,
                 "spconv_fwd_sb: reciprocal of cin_pad %d is inexact at %u", g.cin_pad, kk);
  }
  const int cat = (2 * g.N <= 256) ? 1 : 0;
  int tmem_cols = 32;
  while (tmem_cols < (cat ? 2 * g.N : g.N)) tmem_cols <<= 1;
  tc_launch(spconv_fwd_sb_kernel, tiles, kTcThreads, L.total, stream, (const uint16_t*)features_split,
            (const uint16_t*)packed_sb, pair_fwd, n_out, g.cin_pad, cin_magic, cout, round_up(cout, 8), g.N, kvol,
            g.chunks,
            L.stages, L.stage_bytes, L.pair_off, L.act_off, L.bar_off, tmem_cols, scale, shift, residual, relu, out,
            (uint16_t*)out_split, cat);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

// ---- tile masks + the extended entry (row permutation of mask-sorted tiles, weighted work shares) ----------------
namespace msmd {
// bit k of tile_mask[t]: some row of tile t has a pair at kernel offset k
__global__ void __launch_bounds__(kTcM)
tile_mask_kernel(const int* __restrict__ pair, int kvol, int n_out, uint32_t* __restrict__ tile_mask) {
  __shared__ uint32_t warp_or[kTcM / 32];
  const int o = blockIdx.x * kTcM + threadIdx.x;
  uint32_t m = 0u;
  if (o < n_out)
    for (int k = 0; k < kvol; ++k) m |= (__ldg(pair + (size_t)k * n_out + o) >= 0 ? 1u : 0u) << k;
  m = __reduce_or_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) warp_or[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) tile_mask[blockIdx.x] = warp_or[0] | warp_or[1] | warp_or[2] | warp_or[3];
}
}  // namespace msmd

extern "C" MSMD_API int msmd_rulebook_tile_masks(const int* pair_fwd, int kvol, int n_out, unsigned* tile_mask,
                                                 msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(kvol >= 1 && kvol <= 32 && n_out >= 0, "rulebook_tile_masks: bad sizes");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(pair_fwd && tile_mask, "rulebook_tile_masks: null pointer");
  tile_mask_kernel<<<ceil_div(n_out, kTcM), kTcM, 0, stream>>>(pair_fwd, kvol, n_out, tile_mask);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_fwd_sb_ex(const void* features_split, int n_in, const void* packed_sb,
                                              const int* pair_fwd, const int* row_perm, const unsigned* tile_mask,
                                              int n_out, int cin, int cout, int kvol, const float* scale,
                                              const float* shift, const float* residual, int relu, float* out,
                                              void* out_split, msmd_stream_t stream_) {
  SbGeom g0;
  if (!row_perm && (!tile_mask || !sb_geom(cout, kvol, cin, g0) || !sb_use_persistent(g0)))
    return msmd_spconv_fwd_sb(features_split, n_in, packed_sb, pair_fwd, n_out, cin, cout, kvol, scale, shift,
                              residual, relu, out, out_split, stream_);
  cudaStream_t stream = (cudaStream_t)stream_;
  SbGeom g;
  MSMD_REQUIRE(sb_geom(cout, kvol, cin, g), "spconv_fwd_sb: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0, "spconv_fwd_sb: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd_sb: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(n_in > 0, "spconv_fwd_sb: output rows without input rows");
  MSMD_REQUIRE(features_split && packed_sb && pair_fwd && (out || out_split), "spconv_fwd_sb: null pointer");
  MSMD_REQUIRE(((uintptr_t)features_split & 15) == 0 && ((uintptr_t)packed_sb & 15) == 0 &&
                   ((uintptr_t)out_split & 15) == 0,
               "spconv_fwd_sb: split images and packed weights must be 16-byte aligned");
  return sbp_launch(g, features_split, packed_sb, pair_fwd, row_perm, tile_mask, n_out, cout, kvol, scale, shift,
                    residual, relu, out, out_split, stream);
}
