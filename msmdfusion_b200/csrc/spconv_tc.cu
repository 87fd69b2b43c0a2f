// spconv_tc.cu -- sparse convolution forward on the 5th-generation tensor cores (tcgen05).
//
//   out[o, co] = sum_k sum_ci  x[pair_fwd[k, o], ci] * W[co, k, ci]
//
// is an implicit GEMM  D[M = output voxels, N = Cout] = A[M, K] * B[N, K]^T  with the
// reduction axis K = (kernel offset k, input channel ci):
//   * B is the spconv-2.x KRSC parameter [Cout, kz,ky,kx, Cin] itself (bug_fix/conv.py:114-117),
//     i.e. already K-major.  It is re-packed ONCE per weight into the exact shared-memory image
//     the tensor core reads (32-float K chunks, 128-byte swizzle, split into tf32 hi / lo
//     parts), so a K chunk of B is one contiguous bulk copy through the TMA engine.
//   * A is gathered: row o of chunk j holds x[pair_fwd[k, o], ci..] for the (k, ci) of that
//     chunk.  Producer warps gather 16-byte pieces with coalesced 128-byte row segments, split
//     them into tf32 hi / lo and store them swizzled (the K-major SWIZZLE_128B canonical layout).
//   * One elected thread issues tcgen05.mma.kind::tf32 (M=128, N=Cout, K=8) with the fp32
//     accumulator tile in TMEM.  fp32-level accuracy ("3xTF32"):
//         D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo        (A_lo*B_lo ~ 2^-22 relative, dropped)
//   * Epilogue: tcgen05.ld TMEM -> registers, fused BatchNorm(eval) scale/shift, residual,
//     ReLU, store.
//   K chunks whose kernel offsets no voxel of the tile uses are skipped entirely.
//
// Pipeline: `stages` shared-memory stages, mbarrier full/empty per stage (full = one arrival per
// gather warp + the bulk copy's transaction bytes; empty = tcgen05.commit), one accumulator-ready
// barrier for the epilogue.  Warps 0-7 gather (each thread keeps the loads of TWO chunks in
// flight), then run the epilogue; warp 8 issues the B bulk copies; warp 9 allocates TMEM and
// issues the MMAs.
//
// Variants (chosen by N, see the dispatch at the bottom):
//   variant 2 (this kernel)      A hi/lo staged in shared memory; for 2N <= 256 the B_hi|B_lo halves
//                                are used as ONE 2N-row operand (2 MMAs per k-step instead of 3)
//   variant 3 (spconv_fwd_tc3)   A hi/lo staged in TENSOR MEMORY (tcgen05.st), B only in shared
//                                memory, optional split-K CTA pairs for tail balance
#include "tc_common.cuh"
#include "tc_trace.cuh"

namespace msmd {

struct TcSmemLayout {
  int stage_bytes, stages, pair_off, act_off, bar_off, total;
};

static TcSmemLayout tc_layout(int N, int kvol, int chunks, int tiles) {
  TcSmemLayout L;
  L.stage_bytes = 2 * kTcABytes + N * kTcKC * 4 * 2;
  const int misc = round_up(kvol * kTcM * 4, 16) + round_up(2 * chunks, 16) + 256 + 8 * N;
  // two CTAs per SM when two stages fit in half of the SM's shared memory AND there are enough
  // tiles to need the second slot; otherwise one CTA with a deeper pipeline (max 4 stages)
  const int budget = tc_smem_budget(2 * L.stage_bytes + misc + 1024 <= 112 * 1024, tiles);
  L.stages = (budget - misc - 1024) / L.stage_bytes;
  if (L.stages > tc_stage_cap()) L.stages = tc_stage_cap();
  L.pair_off = L.stages * L.stage_bytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(2 * chunks, 16);
  L.total = L.bar_off + 256 + 8 * N + 1024;  // barriers | scale/shift | slack for the 1024-byte alignment
  return L;
}

template <bool VEC>
__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_tc_kernel(const float* __restrict__ feat, const float* __restrict__ wpk,
                     const int* __restrict__ pair, int n_out, int cin, int cin_pad, int cout, int N,
                     int kvol, int chunks, int stages, int stage_bytes, int pair_off, int act_off,
                     int bar_off, int tmem_cols, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ residual, int relu,
                     float* __restrict__ out, int cat, const int* __restrict__ row_perm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  int* pair_s = (int*)(smem + pair_off);
  unsigned short* alist = (unsigned short*)(smem + act_off);
  uint64_t* full_bar = (uint64_t*)(smem + bar_off);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* accum_bar = full_bar + 8;
  uint32_t* tmem_ptr_s = (uint32_t*)(full_bar + 9);
  int* n_act_s = (int*)(full_bar + 9) + 1;
  int* used_s = (int*)(full_bar + 10);  // [kvol <= 32]
  float* ss = (float*)(smem + bar_off + 256);  // folded BatchNorm scale[N] | shift[N]
  for (int c = threadIdx.x; c < N; c += kTcThreads) {
    ss[c] = (scale && c < cout) ? __ldg(scale + c) : 1.f;
    ss[N + c] = (shift && c < cout) ? __ldg(shift + c) : 0.f;
  }

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcM;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  tc::pdl_launch_dependents();  // (PDL build) the next layer's prologue may overlap this kernel

  // ---- one-time setup -------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], kTcProducerWarps + 1);  // one arrive per gather warp + the B copy's expect_tx
      tc::mbar_init(&empty_bar[s], 1);       // one tcgen05.commit
    }
    tc::mbar_init(accum_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == kTcProducerWarps + 1) {
    tc::tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
    tc::tmem_relinquish();
  }
  // pair table of this tile -> shared memory; which kernel offsets does the tile use?
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
  if (warp == 0) tc_build_active_list(used_s, chunks, cin_pad, kvol, lane, alist, n_act_s);
  __syncthreads();
  const int n_act = *n_act_s;
  const int any_active = n_act > 0;
  const uint32_t tmem_base = *tmem_ptr_s;
  if (tid == 0) { TC_TRACE_HEAD(1, clock64()); TC_TRACE_HEAD(7, n_act); }
  // (PDL build) everything above touched only this layer's constants and its rulebook; features, residual,
  // output and split-K scratch belong to the stream's data flow: wait for the previous kernel here
  tc::pdl_wait();

  if (warp < kTcProducerWarps) {
    // ===== A producers: gather + tf32 hi/lo split + swizzled store ==========================
    // thread -> 16-byte piece p of rows rbase + 32*i: 8 consecutive lanes read one contiguous
    // 128-byte row segment.  The loads of the NEXT active chunk are issued before the current
    // chunk is split and stored, so two chunks (8 x 16 B per thread) are always in flight.
    const int p = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    constexpr int RPT = kTcM / (kTcProducers / 8);  // rows per thread = 4
    auto gather = [&](int j, float4 (&v)[RPT]) {
      const int kk0 = j * kTcKC + p * 4;
      const int k = kk0 / cin_pad;
      const int c = kk0 - k * cin_pad;
      const bool kvalid = k < kvol;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const int idx = kvalid ? pair_s[k * kTcM + r] : -1;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx >= 0) {
          const float* src = feat + (size_t)idx * cin + c;
          if (VEC) {
            v[i] = __ldg((const float4*)src);
          } else {
            if (c + 0 < cin) v[i].x = __ldg(src + 0);
            if (c + 1 < cin) v[i].y = __ldg(src + 1);
            if (c + 2 < cin) v[i].z = __ldg(src + 2);
            if (c + 3 < cin) v[i].w = __ldg(src + 3);
          }
        }
      }
    };
    int it = 0;
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);  // traced gather warps
    (void)tr_role;
    auto store = [&](const float4 (&v)[RPT]) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      if (lane == 0) TC_TRACE(tr_role, it, 0);
      mbar_wait_warp(&empty_bar[s], ph ^ 1u, lane);
      if (lane == 0) TC_TRACE(tr_role, it, 1);
      const uint32_t a_hi = tc::smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t a_lo = a_hi + kTcABytes;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const uint32_t off = (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4));
        const float hx = tc::round_tf32(v[i].x), hy = tc::round_tf32(v[i].y),
                    hz = tc::round_tf32(v[i].z), hw = tc::round_tf32(v[i].w);
        tc::st_shared_v4(a_hi + off, hx, hy, hz, hw);
        tc::st_shared_v4(a_lo + off, v[i].x - hx, v[i].y - hy, v[i].z - hz, v[i].w - hw);
      }
      tc::fence_proxy_async();  // every lane: its generic-proxy stores -> async proxy
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[s]);  // one arrival per warp (256 -> 8 smem atomics)
      if (lane == 0) TC_TRACE(tr_role, it, 2);
      ++it;
    };
    // kTcDepth register buffers rotated without copies: while one chunk is split and stored, the
    // loads of the next kTcDepth-1 chunks are in flight (the gather is L2-latency bound)
    float4 buf[kTcDepth][RPT];
#pragma unroll
    for (int d = 0; d < kTcDepth; ++d)
      if (d < n_act) gather(alist[d], buf[d]);
    for (int i = 0; i < n_act; i += kTcDepth) {
#pragma unroll
      for (int d = 0; d < kTcDepth; ++d) {
        if (i + d < n_act) {
          store(buf[d]);
          if (i + d + kTcDepth < n_act) gather(alist[i + d + kTcDepth], buf[d]);
        }
      }
    }

    // ===== epilogue: TMEM -> registers -> fused BN / residual / ReLU -> global =============
    if (tid == 0) TC_TRACE_HEAD(2, clock64());
    tc_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, N, ss, residual, relu, out,
                nullptr, nullptr, cat ? N : 0, row_perm);
    if (tid == 0) TC_TRACE_HEAD(4, clock64());
  } else if (warp == kTcProducerWarps) {
    // ===== B loader: one bulk copy (hi + lo image of the chunk) per active chunk ============
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)N * kTcKC * 4u * 2u;
      for (int it = 0; it < n_act; ++it) {
        const int j = alist[it];
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        TC_TRACE(2, it, 0);
        tc::mbar_wait(&empty_bar[s], ph ^ 1u);
        TC_TRACE(2, it, 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
        tc::bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kTcABytes,
                     wpk + (size_t)j * N * kTcKC * 2, bytes, &full_bar[s]);
      }
    }
  } else {
    // ===== MMA issuer ========================================================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtTF32, kTcM, N);
      const uint32_t idesc2 = tc::idesc_f32acc(tc::kFmtTF32, kTcM, 2 * N);
      uint32_t accumulate = 0;
      for (int it = 0; it < n_act; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        TC_TRACE(3, it, 0);
        tc::mbar_wait(&full_bar[s], ph);
        TC_TRACE(3, it, 1);
        tc::fence_after_sync();
        const uint32_t a_hi = tc::smem_u32(smem + (size_t)s * stage_bytes);
        const uint32_t a_lo = a_hi + kTcABytes;
        const uint32_t b_hi = a_hi + 2 * kTcABytes;
        const uint32_t b_lo = b_hi + (uint32_t)N * kTcKC * 4u;
#pragma unroll
        for (int ks = 0; ks < kTcKC / 8; ++ks) {
          const uint32_t koff = (uint32_t)ks * 32u;  // 8 tf32 = 32 bytes along K
          const uint64_t dah = tc::desc_k_sw128(a_hi + koff), dal = tc::desc_k_sw128(a_lo + koff);
          const uint64_t dbh = tc::desc_k_sw128(b_hi + koff), dbl = tc::desc_k_sw128(b_lo + koff);
          if (cat) {
            // B_hi and B_lo are adjacent in the stage, i.e. ONE K-major operand of 2N rows:
            //   D[:, 0:2N] += A_hi * [B_hi; B_lo]      D[:, 0:N] += A_lo * B_hi
            // two MMAs per k-step instead of three, A_hi is read from shared memory once; the
            // epilogue adds the two column halves.
            tc::mma_tf32(tmem_base, dah, dbh, idesc2, accumulate);
            tc::mma_tf32(tmem_base, dal, dbh, idesc, 1u);
          } else {
            tc::mma_tf32(tmem_base, dal, dbh, idesc, accumulate);  // small terms first
            tc::mma_tf32(tmem_base, dah, dbl, idesc, 1u);
            tc::mma_tf32(tmem_base, dah, dbh, idesc, 1u);
          }
          accumulate = 1u;
        }
        tc::mma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
        TC_TRACE(3, it, 2);
      }
      if (n_act > 0) tc::mma_commit(accum_bar);  // accumulator complete -> epilogue
    }
  }

  // ---- teardown -------------------------------------------------------------------------
  tc::fence_before_sync();
  __syncthreads();
  TC_TRACE_EXIT();
  if (warp == kTcProducerWarps + 1) tc::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------
// Variant 3: the A operand lives in TENSOR MEMORY.
//
// Variant 2 above is shared-memory-bandwidth bound: per 32-float K chunk it writes A hi+lo
// (32 KB) and the three MMAs of each k-step re-read A from shared memory (48 KB per chunk) on
// top of B (32 KB written, 48 KB read) -- 160 KB per chunk vs 96 KB the SM can move in the 768
// tensor cycles of the chunk (ncu: profiles/r01b_ncu_full_*).  Here the gathered fp32 tile makes ONE
// pass through shared memory (16 KB written coalesced, 16 KB read row-per-thread), is split into
// tf32 hi/lo in registers and stored with tcgen05.st into TMEM (row = lane, K along columns);
// tcgen05.mma reads A from TMEM, so shared memory carries B only.  Shared memory per CTA drops to
// ~110 KB for N = 128, i.e. two CTAs per SM (one CTA's epilogue overlaps the other's main loop).
//
//   smem : B ring (b_stages x N*256 B) | raw A (2 x 16 KB, 128-B swizzled) | pair table | flags | barriers
//   TMEM : D [0, N) | A ring: a_stages x {hi 32 cols | lo 32 cols}
//   barriers: b_full/b_empty (bulk copy <-> MMA), a_full/a_empty (tcgen05.st <-> MMA), accum
// ------------------------------------------------------------------------------------------
struct Tc3Layout {
  int b_stage_bytes, b_stages, a_stages, raw_off, pair_off, act_off, bar_off, total, tmem_cols;
};

static Tc3Layout tc3_layout(int N, int kvol, int chunks) {
  Tc3Layout L;
  L.b_stage_bytes = N * kTcKC * 4 * 2;
  const int misc = 2 * kTcABytes + round_up(kvol * kTcM * 4, 16) + round_up(2 * chunks, 16) + 512 + 8 * N;
  const int half = 113 * 1024, full = 224 * 1024;
  int budget = (2 * L.b_stage_bytes + misc + 1024 <= half) ? half : full;
  if (g_tc_tune[0] == 1) budget = full;   // A/B switch: one CTA per SM
  L.b_stages = (budget - misc - 1024) / L.b_stage_bytes;
  if (L.b_stages > tc_stage_cap()) L.b_stages = tc_stage_cap();
  L.a_stages = (N <= 64) ? 3 : 2;
  L.tmem_cols = 32;
  while (L.tmem_cols < N + 64 * L.a_stages) L.tmem_cols <<= 1;
  L.raw_off = L.b_stages * L.b_stage_bytes;
  L.pair_off = L.raw_off + 2 * kTcABytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(2 * chunks, 16);
  L.total = L.bar_off + 512 + 8 * N + 1024;
  return L;
}

template <bool VEC>
__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_tc3_kernel(const float* __restrict__ feat, const float* __restrict__ wpk,
                      const int* __restrict__ pair, int n_out, int cin, int cin_pad, int cout, int N,
                      int kvol, int chunks, int b_stages, int b_stage_bytes, int a_stages, int raw_off,
                      int pair_off, int act_off, int bar_off, int tmem_cols,
                      const float* __restrict__ scale, const float* __restrict__ shift,
                      const float* __restrict__ residual, int relu, float* __restrict__ out, int split,
                      float* __restrict__ part_ws, int* __restrict__ part_flag,
                      const int* __restrict__ row_perm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  int* pair_s = (int*)(smem + pair_off);
  const unsigned short* alist = (unsigned short*)(smem + act_off);
  uint64_t* b_full = (uint64_t*)(smem + bar_off);
  uint64_t* b_empty = b_full + 4;
  uint64_t* a_full = b_full + 8;
  uint64_t* a_empty = b_full + 12;
  uint64_t* accum_bar = b_full + 16;
  uint32_t* tmem_ptr_s = (uint32_t*)(b_full + 17);
  int* n_act_s = (int*)(b_full + 17) + 1;
  int* used_s = (int*)(b_full + 18);  // [kvol <= 32]
  float* ss = (float*)(smem + bar_off + 512);  // folded BatchNorm scale[N] | shift[N]
  for (int c = threadIdx.x; c < N; c += kTcThreads) {
    ss[c] = (scale && c < cout) ? __ldg(scale + c) : 1.f;
    ss[N + c] = (shift && c < cout) ? __ldg(shift + c) : 0.f;
  }

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // split-K pairs (tail balance): CTA 2t and 2t+1 share output tile t and take half of its
  // active K chunks each; the first to finish hands its partial sums to the second through L2
  const int tile = (int)blockIdx.x / split, half = (int)blockIdx.x % split;
  const int row0 = tile * kTcM;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  tc::pdl_launch_dependents();  // (PDL build) the next layer's prologue may overlap this kernel

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&b_full[s], 1);              // the arrive.expect_tx of the bulk copy
      tc::mbar_init(&b_empty[s], 1);             // tcgen05.commit
      tc::mbar_init(&a_full[s], kTcProducerWarps);  // one arrive per converter warp after its tcgen05.st
      tc::mbar_init(&a_empty[s], 1);             // tcgen05.commit
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
  if (warp == 0)
    tc_build_active_list(used_s, chunks, cin_pad, kvol, lane, (unsigned short*)alist, n_act_s);
  __syncthreads();
  int n_act = *n_act_s;
  if (split == 2) {
    const int mid = n_act / 2;
    if (half) { alist += mid; n_act -= mid; } else { n_act = mid; }
  }
  const int any_active = n_act > 0;
  const uint32_t tmem_base = *tmem_ptr_s;
  const uint32_t tmem_a0 = tmem_base + (uint32_t)N;  // A ring starts right after the accumulator
  if (tid == 0) { TC_TRACE_HEAD(1, clock64()); TC_TRACE_HEAD(7, n_act); }
  // (PDL build) everything above touched only this layer's constants and its rulebook; features, residual,
  // output and split-K scratch belong to the stream's data flow: wait for the previous kernel here
  tc::pdl_wait();

  if (warp < kTcProducerWarps) {
    // ===== gather (coalesced) -> raw smem -> row-per-thread read -> split -> tcgen05.st =====
    const int p = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    constexpr int RPT = kTcM / (kTcProducers / 8);  // 4 rows per thread in the gather mapping
    auto gather = [&](int j, float4 (&v)[RPT]) {
      const int kk0 = j * kTcKC + p * 4;
      const int k = kk0 / cin_pad;
      const int c = kk0 - k * cin_pad;
      const bool kvalid = k < kvol;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const int idx = kvalid ? pair_s[k * kTcM + r] : -1;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx >= 0) {
          const float* src = feat + (size_t)idx * cin + c;
          if (VEC) {
            v[i] = __ldg((const float4*)src);
          } else {
            if (c + 0 < cin) v[i].x = __ldg(src + 0);
            if (c + 1 < cin) v[i].y = __ldg(src + 1);
            if (c + 2 < cin) v[i].z = __ldg(src + 2);
            if (c + 3 < cin) v[i].w = __ldg(src + 3);
          }
        }
      }
    };
    // converter mapping: this thread owns accumulator row `crow` (the TMEM lane it may write)
    // and 16 of the chunk's 32 K columns
    const int crow = (warp & 3) * 32 + lane;
    const int chalf = warp >> 2;  // K columns [16*chalf, +16)
    const uint32_t raw0 = tc::smem_u32(smem + raw_off);
    int it = 0;
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);  // traced gather warps
    (void)tr_role;
    auto convert = [&](const float4 (&v)[RPT]) {
      const uint32_t raw = raw0 + (uint32_t)(it & 1) * kTcABytes;
      // 1. coalesced-layout registers -> swizzled raw tile (one pass, conflict-free)
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        tc::st_shared_v4(raw + (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)), v[i].x, v[i].y, v[i].z, v[i].w);
      }
      tc::named_bar_sync(1, kTcProducers);
      // 2. my row, my 16 K columns (4 pieces), back out of the raw tile
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int piece = chalf * 4 + q;
        const float4 x = tc::ld_shared_v4(raw + (uint32_t)(crow * 128 + ((piece ^ (crow & 7)) << 4)));
        const float hx = tc::round_tf32(x.x), hy = tc::round_tf32(x.y), hz = tc::round_tf32(x.z),
                    hw = tc::round_tf32(x.w);
        hi[4 * q + 0] = __float_as_uint(hx); hi[4 * q + 1] = __float_as_uint(hy);
        hi[4 * q + 2] = __float_as_uint(hz); hi[4 * q + 3] = __float_as_uint(hw);
        lo[4 * q + 0] = __float_as_uint(x.x - hx); lo[4 * q + 1] = __float_as_uint(x.y - hy);
        lo[4 * q + 2] = __float_as_uint(x.z - hz); lo[4 * q + 3] = __float_as_uint(x.w - hw);
      }
      // 3. TMEM A stage free?  store hi | lo, make it visible to the MMA issuer
      const int sa = it % a_stages;
      const uint32_t ph = (uint32_t)(it / a_stages) & 1u;
      if (lane == 0) TC_TRACE(tr_role, it, 0);
      mbar_wait_warp(&a_empty[sa], ph ^ 1u, lane);
      if (lane == 0) TC_TRACE(tr_role, it, 1);
      tc::fence_after_sync();
      const uint32_t ta = tmem_a0 + (uint32_t)(sa * 64) + ((uint32_t)((warp & 3) * 32) << 16);
      tc::tmem_st16(ta + (uint32_t)(16 * chalf), hi);
      tc::tmem_st16(ta + 32u + (uint32_t)(16 * chalf), lo);
      tc::tmem_st_wait();  // warp-wide: all 32 lanes' stores have landed
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&a_full[sa]);
      if (lane == 0) TC_TRACE(tr_role, it, 2);
      ++it;
    };
    float4 bufa[RPT], bufb[RPT];
    if (n_act > 0) gather(alist[0], bufa);
    if (n_act > 1) gather(alist[1], bufb);
    for (int i = 0; i < n_act; i += 2) {
      convert(bufa);
      if (i + 2 < n_act) gather(alist[i + 2], bufa);
      if (i + 1 >= n_act) break;
      convert(bufb);
      if (i + 3 < n_act) gather(alist[i + 3], bufb);
    }

    if (tid == 0) TC_TRACE_HEAD(2, clock64());
    if (split == 1) {
      tc_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, N, ss, residual, relu, out,
                  nullptr, nullptr, 0, row_perm);
      if (tid == 0) TC_TRACE_HEAD(4, clock64());
    } else {
      int* ticket_s = n_act_s;  // the active-chunk count is no longer needed: reuse its smem word
      float* part = part_ws + (size_t)tile * kTcM * N;
      tc::named_bar_sync(2, kTcProducers);  // every thread has read *n_act_s
      if (tid == 0) *ticket_s = atomicAdd(&part_flag[2 * tile], 1);
      tc::named_bar_sync(2, kTcProducers);
      if (*ticket_s == 0) {
        tc_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, N, ss, residual, relu,
                    out, part, nullptr);
        __threadfence();
        tc::named_bar_sync(2, kTcProducers);
        if (tid == 0) atomicExch(&part_flag[2 * tile + 1], 1);  // partial sums are in L2
      } else {
        if (tid == 0) {
          const long long t0 = clock64();
          while (atomicAdd(&part_flag[2 * tile + 1], 0) == 0) {
            if (clock64() - t0 > 4000000000LL) __trap();  // the partner is already in its epilogue
          }
        }
        tc::named_bar_sync(2, kTcProducers);
        __threadfence();
        tc_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, N, ss, residual, relu,
                    out, nullptr, part, 0, row_perm);
        tc::named_bar_sync(2, kTcProducers);
        if (tid == 0) { part_flag[2 * tile] = 0; part_flag[2 * tile + 1] = 0; }  // ready for the next launch
      }
    }
  } else if (warp == kTcProducerWarps) {
    // ===== B loader ============================================================================
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)b_stage_bytes;
      for (int it = 0; it < n_act; ++it) {
        const int j = alist[it];
        const int s = it % b_stages;
        const uint32_t ph = (uint32_t)(it / b_stages) & 1u;
        TC_TRACE(2, it, 0);
        tc::mbar_wait(&b_empty[s], ph ^ 1u);
        TC_TRACE(2, it, 1);
        tc::mbar_arrive_expect_tx(&b_full[s], bytes);
        tc::bulk_g2s(smem + (size_t)s * b_stage_bytes, wpk + (size_t)j * N * kTcKC * 2, bytes, &b_full[s]);
      }
    }
  } else {
    // ===== MMA issuer: A from TMEM, B from shared memory =======================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtTF32, kTcM, N);
      uint32_t accumulate = 0;
      for (int it = 0; it < n_act; ++it) {
        const int sb = it % b_stages, sa = it % a_stages;
        TC_TRACE(3, it, 0);
        tc::mbar_wait(&b_full[sb], (uint32_t)(it / b_stages) & 1u);
        TC_TRACE(3, it, 3);   // weights in; now the A operand
        tc::mbar_wait(&a_full[sa], (uint32_t)(it / a_stages) & 1u);
        TC_TRACE(3, it, 1);
        tc::fence_after_sync();
        const uint32_t b_hi = tc::smem_u32(smem + (size_t)sb * b_stage_bytes);
        const uint32_t b_lo = b_hi + (uint32_t)N * kTcKC * 4u;
        const uint32_t a_hi = tmem_a0 + (uint32_t)(sa * 64);
        const uint32_t a_lo = a_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < kTcKC / 8; ++ks) {
          const uint64_t dbh = tc::desc_k_sw128(b_hi + (uint32_t)ks * 32u);
          const uint64_t dbl = tc::desc_k_sw128(b_lo + (uint32_t)ks * 32u);
          tc::mma_tf32_ts(tmem_base, a_lo + (uint32_t)ks * 8u, dbh, idesc, accumulate);
          tc::mma_tf32_ts(tmem_base, a_hi + (uint32_t)ks * 8u, dbl, idesc, 1u);
          tc::mma_tf32_ts(tmem_base, a_hi + (uint32_t)ks * 8u, dbh, idesc, 1u);
          accumulate = 1u;
        }
        tc::mma_commit(&b_empty[sb]);
        tc::mma_commit(&a_empty[sa]);
        TC_TRACE(3, it, 2);
      }
      if (n_act > 0) tc::mma_commit(accum_bar);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  TC_TRACE_EXIT();
  if (warp == kTcProducerWarps + 1) tc::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// Packed weight image: [chunk j][half: hi, lo][n < N][32 floats, 128-byte swizzled by (n & 7)].
__global__ void __launch_bounds__(256)
tc_pack_weight_kernel(const float* __restrict__ w, int cout, int kvol, int cin, int cin_pad, int N,
                      int chunks, float* __restrict__ packed) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)chunks * N * kTcKC;
  if (t >= total) return;
  const int kk = (int)(t % kTcKC);
  const int n = (int)((t / kTcKC) % N);
  const int j = (int)(t / ((size_t)kTcKC * N));
  const int K = j * kTcKC + kk;
  const int k = K / cin_pad, c = K - k * cin_pad;
  float val = 0.f;
  if (k < kvol && c < cin && n < cout) val = w[((size_t)n * kvol + k) * cin + c];
  const float hi = tc::round_tf32(val);
  const float lo = val - hi;
  const size_t blk = (size_t)N * kTcKC;
  const size_t off = (size_t)n * kTcKC + (size_t)((((kk >> 2) ^ (n & 7)) << 2) + (kk & 3));
  packed[((size_t)j * 2 + 0) * blk + off] = hi;
  packed[((size_t)j * 2 + 1) * blk + off] = lo;
}

struct TcGeom {
  int cin_pad, N, chunks, tmem_cols;
};
static bool tc_geom(int cout, int kvol, int cin, TcGeom& g) {
  if (cout < 1 || cout > 256 || kvol < 1 || kvol > 32 || cin < 1) return false;
  g.cin_pad = round_up(cin, 4);
  g.N = round_up(cout, 16);
  g.chunks = ((long long)kvol * g.cin_pad + kTcKC - 1) / kTcKC;
  g.tmem_cols = 32;
  while (g.tmem_cols < g.N) g.tmem_cols <<= 1;
  return true;
}

}  // namespace msmd

using namespace msmd;

#ifdef MSMD_TC_TRACE
extern "C" MSMD_API int msmd_tc_trace_set(unsigned long long* buf) { return tc_trace_set_impl(buf); }
extern "C" MSMD_API int msmd_tc_trace_record_words(void) { return kTrRecord; }
#endif

int msmd::g_tc_tune[5] = {0, 0, 0, 0, 0};

extern "C" MSMD_API int msmd_spconv_tc_set_tuning(int key, int value) {
  MSMD_REQUIRE(key >= 0 && key < 5, "spconv_tc_set_tuning: key must be 0 (occupancy), 1 (stage cap), 2 (split-K), 3 (chunks per stage) or 4 (epilogue units)");
  g_tc_tune[key] = value;
  return MSMD_OK;
}

static int g_tc_variant = 0;  // 0: auto (by N); 2: A through shared memory; 3: A through tensor memory

extern "C" MSMD_API int msmd_spconv_tc_set_variant(int variant) {
  MSMD_REQUIRE(variant == 0 || variant == 2 || variant == 3, "spconv_tc_set_variant: variant must be 0, 2 or 3");
  g_tc_variant = variant;
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_tc_supported(int cout, int kvol, int cin) {
  TcGeom g;
  return tc_geom(cout, kvol, cin, g) ? 1 : 0;
}

extern "C" MSMD_API size_t msmd_spconv_tc_packed_floats(int cout, int kvol, int cin) {
  TcGeom g;
  if (!tc_geom(cout, kvol, cin, g)) return 0;
  return (size_t)g.chunks * 2 * g.N * kTcKC;
}

extern "C" MSMD_API int msmd_spconv_tc_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                                   float* packed, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcGeom g;
  MSMD_REQUIRE(tc_geom(cout, kvol, cin, g), "spconv_tc: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(weight_krsc && packed, "spconv_tc_pack_weight: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed & 15) == 0, "spconv_tc_pack_weight: packed must be 16-byte aligned");
  const size_t total = (size_t)g.chunks * g.N * kTcKC;
  tc_pack_weight_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(
      weight_krsc, cout, kvol, cin, g.cin_pad, g.N, g.chunks, packed);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

// Split-K pairs pay off when the tile count leaves most SMs idle or one tile short of a wave:
// makespan in tile-times is ceil(t/148) unsplit and ceil(2t/148)/2 split.
static bool tc_use_split(int tiles, int N) {
  if (N < 96 || g_tc_tune[2] == 1) return false;  // variant 3 only
  if (g_tc_tune[2] == 2) return true;
  return tiles <= kNumSMs / 2 || (tiles > kNumSMs && tiles <= kNumSMs + kNumSMs / 2);
}

extern "C" MSMD_API size_t msmd_spconv_tc_workspace(int n_out, int cout) {
  const int N = round_up(cout > 0 ? cout : 1, 16);
  const int tiles = ceil_div(n_out > 0 ? n_out : 1, kTcM);
  if (g_tc_variant == 2 || !tc_use_split(tiles, N)) return 0;
  return (size_t)tiles * kTcM * N * sizeof(float) + (size_t)tiles * 2 * sizeof(int) + 512;
}

extern "C" MSMD_API int msmd_spconv_fwd_tc(const float* features, int n_in, const float* packed_tc,
                                           const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                           const float* scale, const float* shift,
                                           const float* residual, int relu, float* out,
                                           msmd_stream_t stream_) {
  return msmd_spconv_fwd_tc_ws(features, n_in, packed_tc, pair_fwd, n_out, cin, cout, kvol, scale, shift,
                               residual, relu, out, nullptr, 0, stream_);
}

static int tc_forward_impl(const float* features, int n_in, const float* packed_tc, const int* pair_fwd,
                           const int* row_perm, int n_out, int cin, int cout, int kvol, const float* scale,
                           const float* shift, const float* residual, int relu, float* out, void* workspace,
                           size_t workspace_bytes, msmd_stream_t stream_);

extern "C" MSMD_API int msmd_spconv_fwd_tc_ws(const float* features, int n_in, const float* packed_tc,
                                           const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                           const float* scale, const float* shift,
                                           const float* residual, int relu, float* out,
                                           void* workspace, size_t workspace_bytes,
                                           msmd_stream_t stream_) {
  return tc_forward_impl(features, n_in, packed_tc, pair_fwd, nullptr, n_out, cin, cout, kvol, scale, shift,
                         residual, relu, out, workspace, workspace_bytes, stream_);
}

extern "C" MSMD_API int msmd_spconv_fwd_tc_sorted(const float* features, int n_in, const float* packed_tc,
                                                  const int* pair_sorted, const int* row_perm, int n_out,
                                                  int cin, int cout, int kvol, const float* scale,
                                                  const float* shift, const float* residual, int relu,
                                                  float* out, void* workspace, size_t workspace_bytes,
                                                  msmd_stream_t stream_) {
  MSMD_REQUIRE(row_perm || n_out == 0, "spconv_fwd_tc_sorted: null row_perm");
  return tc_forward_impl(features, n_in, packed_tc, pair_sorted, row_perm, n_out, cin, cout, kvol, scale,
                         shift, residual, relu, out, workspace, workspace_bytes, stream_);
}

static int tc_forward_impl(const float* features, int n_in, const float* packed_tc, const int* pair_fwd,
                           const int* row_perm, int n_out, int cin, int cout, int kvol, const float* scale,
                           const float* shift, const float* residual, int relu, float* out, void* workspace,
                           size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcGeom g;
  MSMD_REQUIRE(tc_geom(cout, kvol, cin, g), "spconv_fwd_tc: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0, "spconv_fwd_tc: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd_tc: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(features && packed_tc && pair_fwd && out, "spconv_fwd_tc: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed_tc & 15) == 0, "spconv_fwd_tc: packed weights must be 16-byte aligned");
  const bool vec = (cin % 4 == 0) && (((uintptr_t)features & 15) == 0);
  const int tiles = ceil_div(n_out, kTcM);
  static bool attr_set[4] = {false, false, false, false};
  // variant 3 (A in tensor memory) wins once the tile is shared-memory-bandwidth bound (N >= 96);
  // below that the per-chunk latency chain of variant 2 is shorter (measured, profiles/README.md)
  if (g_tc_variant == 3 || (g_tc_variant == 0 && g.N >= 96)) {
    const Tc3Layout L = tc3_layout(g.N, kvol, g.chunks);
    MSMD_REQUIRE(L.b_stages >= 1 && L.tmem_cols <= 512, "spconv_fwd_tc: tile does not fit on the SM");
    auto kern = vec ? spconv_fwd_tc3_kernel<true> : spconv_fwd_tc3_kernel<false>;
    if (!attr_set[2 + vec]) {
      MSMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set[2 + vec] = true;
    }
    int split = 1;
    float* part_ws = nullptr;
    int* part_flag = nullptr;
    const size_t need = msmd_spconv_tc_workspace(n_out, cout);
    if (workspace && need > 0 && workspace_bytes >= need) {
      split = 2;
      part_flag = (int*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
      part_ws = (float*)(part_flag + (size_t)2 * tiles + (64 - (2 * tiles) % 64) % 64);
      MSMD_CUDA_OK(cudaMemsetAsync(part_flag, 0, (size_t)2 * tiles * sizeof(int), stream));
    }
    tc_launch(kern, tiles * split, kTcThreads, L.total, stream, 
        features, packed_tc, pair_fwd, n_out, cin, g.cin_pad, cout, g.N, kvol, g.chunks, L.b_stages,
        L.b_stage_bytes, L.a_stages, L.raw_off, L.pair_off, L.act_off, L.bar_off, L.tmem_cols, scale, shift,
        residual, relu, out, split, part_ws, part_flag, row_perm);
    MSMD_LAUNCH_OK();
    return MSMD_OK;
  }
  const TcSmemLayout L = tc_layout(g.N, kvol, g.chunks, tiles);
  MSMD_REQUIRE(L.stages >= 2, "spconv_fwd_tc: tile does not fit in shared memory");
  auto kern = vec ? spconv_fwd_tc_kernel<true> : spconv_fwd_tc_kernel<false>;
  if (!attr_set[vec]) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[vec] = true;
  }
  // concatenated-B mode needs 2N accumulator columns; two co-resident CTAs must fit in 512
  const int cat = (2 * g.N <= 256) ? 1 : 0;
  int tmem_cols = g.tmem_cols;
  if (cat) {
    tmem_cols = 32;
    while (tmem_cols < 2 * g.N) tmem_cols <<= 1;
  }
  tc_launch(kern, tiles, kTcThreads, L.total, stream, features, packed_tc, pair_fwd, n_out, cin, g.cin_pad, cout,
                                               g.N, kvol, g.chunks, L.stages, L.stage_bytes, L.pair_off,
                                               L.act_off, L.bar_off, tmem_cols, scale, shift, residual,
                                               relu, out, cat, row_perm);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
