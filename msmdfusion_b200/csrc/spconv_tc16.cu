// spconv_tc16.cu -- sparse convolution forward on tcgen05 with 16-bit (bf16) operands.
//
// Same implicit GEMM as spconv_tc.cu (D[128 output voxels, Cout] = A[gathered rows, K] * B[Cout, K]^T,
// K = (kernel offset, input channel), fp32 accumulator in tensor memory, fused BN / residual / ReLU
// epilogue), with tcgen05.mma.kind::f16 on bf16 operands -- twice the tensor rate of kind::tf32 and twice
// the K elements per 128-byte shared-memory row, so a K chunk is 64 elements and a tile runs HALF the
// pipeline steps of the tf32 kernel.  Two modes (msmd_conv_layer.weight_tc / the `x3` argument):
//
//   bf16      one MMA per product.  Operands rounded to bf16 (2^-9 relative): the train-step arithmetic
//             of BASELINE configs[4] ("bf16 sparse-conv kernels", fp32 accumulate, fp32 master weights);
//             NOT within the 1e-4 inference parity bound.
//   bf16 x3   x = hi + lo with hi = bf16(x), lo = bf16(x - hi) (16 significand bits kept);
//             D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, the lo*lo term (2^-18) dropped: ~5e-6 relative per
//             layer (tests/test_oracle.py::test_bf16x3_accuracy_model), i.e. inside the parity bound at
//             half the tensor-pipe time of 3xTF32.  Opt-in until measured on hardware.
//
// Features stay fp32 in HBM: the gather warps load float4 pieces exactly like the tf32 kernel (32 K
// elements per step, kT16Depth steps in flight per thread), convert to bf16 in registers and fill one
// half of the 128-byte row per step, so the register budget is the tf32 kernel's.  The operand layout in
// shared memory is the K-major SWIZZLE_128B canonical layout the tf32 kernel uses (same descriptors; only
// the instruction descriptor's operand format and K = 16 per MMA differ).
//
// Status: written without GPU time; then confirmed on a B200 by tools/quick_gpu_check.py (correctness at four
// channel widths, profiles/r01h_quick_gpu_check.json) -- not yet timed.  Checked on the host model of tcgen05 (tests/tools/cuda_emul/tc_emul.h,
// calibrated on the GPU-verified tf32 kernels): tests/test_cuda_emul.py::test_tc16_*.  GPU tests:
// tests/test_zz_train_gpu.py::test_tc16_*.
#include "tc_common.cuh"
#include "tc_trace.cuh"

namespace msmd {

constexpr int kT16KC = 64;                 // bf16 K elements per chunk = one 128-byte swizzle row
constexpr int kT16Step = 32;               // K elements per gather step (one float4 per thread and row)
constexpr int kT16ABytes = kTcM * 128;     // 16 KB per A image (hi or lo)
constexpr int kT16Depth = 3;               // gather steps whose loads are in flight per thread

struct T16Layout {
  int stage_bytes, stages, pair_off, act_off, bar_off, total;
};

static T16Layout t16_layout(int N, int kvol, int chunks, int tiles, int images, int cps = 1) {
  T16Layout L;
  // one chunk block = A hi [| A lo] | B hi [| B lo], all 1024-B aligned; a stage holds `cps` of them
  L.stage_bytes = cps * images * (kT16ABytes + N * 128);
  const int misc = round_up(kvol * kTcM * 4, 16) + round_up(2 * chunks, 16) + 256 + 8 * N;
  const int budget = tc_smem_budget(2 * L.stage_bytes + misc + 1024 <= 112 * 1024, tiles);  // half: two CTAs per SM
  L.stages = (budget - misc - 1024) / L.stage_bytes;
  if (L.stages > tc_stage_cap()) L.stages = tc_stage_cap();
  L.pair_off = L.stages * L.stage_bytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(2 * chunks, 16);
  L.total = L.bar_off + 256 + 8 * N + 1024;
  return L;
}

template <bool VEC, bool X3>
__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_tc16_kernel(const float* __restrict__ feat, const uint16_t* __restrict__ wpk,
                       const int* __restrict__ pair, int n_out, int cin, int cin_pad, int cout, int N,
                       int kvol, int chunks, int stages, int stage_bytes, int cps, int pair_off, int act_off,
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
  constexpr int kImages = X3 ? 2 : 1;
  // a pipeline stage holds `cps` (1 or 2) chunk blocks: one mbarrier round trip per cps * 64 K elements
  const int chunk_bytes = stage_bytes / cps;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcM;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  tc::pdl_launch_dependents();  // (PDL build) the next layer's prologue may overlap this kernel

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], kTcProducerWarps + 1);  // one arrive per gather warp + the B copy's expect_tx
      tc::mbar_init(&empty_bar[s], 1);                    // one tcgen05.commit
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
  if (warp == 0) tc_build_active_list(used_s, chunks, cin_pad, kvol, lane, alist, n_act_s, kT16KC);
  __syncthreads();
  const int n_act = *n_act_s;
  const int any_active = n_act > 0;
  const uint32_t tmem_base = *tmem_ptr_s;
  if (tid == 0) { TC_TRACE_HEAD(1, clock64()); TC_TRACE_HEAD(7, n_act); }
  // (PDL build) everything above touched only this layer's constants and its rulebook; features, residual,
  // output and split-K scratch belong to the stream's data flow: wait for the previous kernel here
  tc::pdl_wait();

  if (warp < kTcProducerWarps) {
    // ===== A producers: gather (fp32) -> bf16 [hi | lo] -> swizzled store ======================
    // A chunk is filled in two gather steps of 32 K elements; in a step thread (p, rbase) owns
    // the float4 piece p of rows rbase + 32*i, i.e. 8 bytes of the bf16 row: the lower (p even)
    // or upper (p odd) half of the 16-byte swizzle unit  4*h + p/2  of the row.
    const int p = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    constexpr int RPT = kTcM / (kTcProducers / 8);  // rows per thread = 4
    auto gather = [&](int t, float4 (&v)[RPT]) {   // t = gather step: chunk alist[t / 2], half t % 2
      const int kk0 = ((int)alist[t >> 1] * 2 + (t & 1)) * kT16Step + p * 4;
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
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);  // traced gather warps
    (void)tr_role;
    auto store = [&](int t, const float4 (&v)[RPT]) {
      const int ci = t >> 1, h = t & 1;       // chunk of the active list, half of the chunk
      const int it = ci / cps, c = ci - it * cps;  // stage use (group of cps chunks), slot inside the stage
      const int s = it % stages;
      if (h == 0 && c == 0) {
        if (lane == 0) TC_TRACE(tr_role, it, 0);
        mbar_wait_warp(&empty_bar[s], ((uint32_t)(it / stages) & 1u) ^ 1u, lane);
        if (lane == 0) TC_TRACE(tr_role, it, 1);
      }
      const uint32_t a_hi = tc::smem_u32(smem + (size_t)s * stage_bytes + (size_t)c * chunk_bytes);
      const uint32_t a_lo = a_hi + kT16ABytes;
      const int unit = 4 * h + (p >> 1);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        const uint32_t off = (uint32_t)(r * 128 + ((unit ^ (r & 7)) << 4) + ((p & 1) << 3));
        const uint32_t h01 = tc::pack_bf16x2(v[i].x, v[i].y), h23 = tc::pack_bf16x2(v[i].z, v[i].w);
        tc::st_shared_v2_b32(a_hi + off, h01, h23);
        if (X3) {
          const float lx = v[i].x - __uint_as_float(h01 << 16), ly = v[i].y - __uint_as_float(h01 & 0xFFFF0000u);
          const float lz = v[i].z - __uint_as_float(h23 << 16), lw = v[i].w - __uint_as_float(h23 & 0xFFFF0000u);
          tc::st_shared_v2_b32(a_lo + off, tc::pack_bf16x2(lx, ly), tc::pack_bf16x2(lz, lw));
        }
      }
      if (h == 1 && (c == cps - 1 || ci == n_act - 1)) {  // the stage's chunks are complete
        tc::fence_proxy_async();  // every lane: its generic-proxy stores -> async proxy
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full_bar[s]);
        if (lane == 0) TC_TRACE(tr_role, it, 2);
      }
    };
    const int n_steps = 2 * n_act;
    float4 buf[kT16Depth][RPT];
#pragma unroll
    for (int d = 0; d < kT16Depth; ++d)
      if (d < n_steps) gather(d, buf[d]);
    for (int i = 0; i < n_steps; i += kT16Depth) {
#pragma unroll
      for (int d = 0; d < kT16Depth; ++d) {
        if (i + d < n_steps) {
          store(i + d, buf[d]);
          if (i + d + kT16Depth < n_steps) gather(i + d + kT16Depth, buf[d]);
        }
      }
    }

    if (tid == 0) TC_TRACE_HEAD(2, clock64());
    tc_epilogue(any_active, accum_bar, tmem_base, warp, lane, row0, n_out, cout, N, ss, residual, relu, out,
                nullptr, nullptr, (X3 && cat) ? N : 0, row_perm);
    if (tid == 0) TC_TRACE_HEAD(4, clock64());
  } else if (warp == kTcProducerWarps) {
    // ===== B loader: one bulk copy (hi [+ lo] image of the chunk) per active chunk ==============
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)(kImages * N * 128);
      const int n_groups = (n_act + cps - 1) / cps;
      for (int it = 0; it < n_groups; ++it) {
        const int cnt = min(cps, n_act - it * cps);
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        TC_TRACE(2, it, 0);
        tc::mbar_wait(&empty_bar[s], ph ^ 1u);
        TC_TRACE(2, it, 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], bytes * (uint32_t)cnt);
        for (int c = 0; c < cnt; ++c)
          tc::bulk_g2s(smem + (size_t)s * stage_bytes + (size_t)c * chunk_bytes + kImages * kT16ABytes,
                       (const uint8_t*)wpk + (size_t)alist[it * cps + c] * bytes, bytes, &full_bar[s]);
      }
    }
  } else {
    // ===== MMA issuer ===========================================================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtBF16, kTcM, N);
      const uint32_t idesc2 = tc::idesc_f32acc(tc::kFmtBF16, kTcM, 2 * N);
      uint32_t accumulate = 0;
      const int n_groups = (n_act + cps - 1) / cps;
      for (int it = 0; it < n_groups; ++it) {
        const int cnt = min(cps, n_act - it * cps);
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        TC_TRACE(3, it, 0);
        tc::mbar_wait(&full_bar[s], ph);
        TC_TRACE(3, it, 1);
        tc::fence_after_sync();
        for (int c = 0; c < cnt; ++c) {
          const uint32_t a_hi = tc::smem_u32(smem + (size_t)s * stage_bytes + (size_t)c * chunk_bytes);
          const uint32_t a_lo = a_hi + kT16ABytes;
          const uint32_t b_hi = a_hi + kImages * kT16ABytes;
          const uint32_t b_lo = b_hi + (uint32_t)N * 128u;
#pragma unroll
          for (int ks = 0; ks < kT16KC / 16; ++ks) {
            const uint32_t koff = (uint32_t)ks * 32u;  // 16 bf16 = 32 bytes along K
            const uint64_t dah = tc::desc_k_sw128(a_hi + koff), dbh = tc::desc_k_sw128(b_hi + koff);
            if (!X3) {
              tc::mma_f16(tmem_base, dah, dbh, idesc, accumulate);
            } else {
              const uint64_t dal = tc::desc_k_sw128(a_lo + koff), dbl = tc::desc_k_sw128(b_lo + koff);
              if (cat) {
                // B_hi and B_lo are adjacent in the chunk block = ONE K-major operand of 2N rows:
                //   D[:, 0:2N] += A_hi * [B_hi; B_lo]      D[:, 0:N] += A_lo * B_hi
                tc::mma_f16(tmem_base, dah, dbh, idesc2, accumulate);
                tc::mma_f16(tmem_base, dal, dbh, idesc, 1u);
              } else {
                tc::mma_f16(tmem_base, dal, dbh, idesc, accumulate);  // small terms first
                tc::mma_f16(tmem_base, dah, dbl, idesc, 1u);
                tc::mma_f16(tmem_base, dah, dbh, idesc, 1u);
              }
            }
            accumulate = 1u;
          }
        }
        tc::mma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
        TC_TRACE(3, it, 2);
      }
      if (n_act > 0) tc::mma_commit(accum_bar);  // accumulator complete -> epilogue
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  TC_TRACE_EXIT();
  if (warp == kTcProducerWarps + 1) tc::tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

// ------------------------------------------------------------------------------------------
// Variant T: the A operand lives in TENSOR MEMORY (the 16-bit counterpart of spconv_fwd_tc3_kernel).
//
// The kernel above keeps A hi [+ lo] in shared memory: a stage is 2 x (16 KB + N*128 B) in the x3 mode, so at
// N >= 96 only one CTA fits on an SM.  Here the gathered fp32 half-chunk (32 K elements) makes one pass through
// a 16 KB swizzled staging tile, thread = row reads 16 of its 32 values back, converts them to 8 packed bf16x2
// words (hi) [+ 8 (lo)] and writes them with tcgen05.st.x8 into the chunk's TMEM stage -- 32 columns hold the 64 K
// elements of a chunk (two per column; element 2c in the low half of column c), hi at [0, 32), lo at [32, 64).
// tcgen05.mma reads A from TMEM (8 columns per K = 16 step), shared memory carries only the weight ring:
// ~110 KB at N = 128, i.e. two CTAs per SM, plus the split-K pairs of the tf32 kernel for tail balance.
//   smem : B ring (b_stages x images*N*128 B) | raw A (2 x 16 KB) | pair table | flags | barriers
//   TMEM : D [0, N) | A ring: a_stages x {hi 32 cols | lo 32 cols}
// Opt-in inside the opt-in (msmd_spconv_tc16_set_variant / MSMD_TC16_VARIANT=3): the 16-bit TMEM operand layout
// is taken from the PTX ISA text and has not been exercised on hardware by this project.
// ------------------------------------------------------------------------------------------
struct T16tLayout {
  int b_stage_bytes, b_stages, a_stages, raw_off, pair_off, act_off, bar_off, total, tmem_cols;
};

static T16tLayout t16t_layout(int N, int kvol, int chunks, int images) {
  T16tLayout L;
  L.b_stage_bytes = images * N * 128;
  const int misc = 2 * kT16ABytes + round_up(kvol * kTcM * 4, 16) + round_up(2 * chunks, 16) + 512 + 8 * N;
  const int half = 113 * 1024, full = 224 * 1024;
  int budget = (2 * L.b_stage_bytes + misc + 1024 <= half) ? half : full;
  if (g_tc_tune[0] == 1) budget = full;
  L.b_stages = (budget - misc - 1024) / L.b_stage_bytes;
  if (L.b_stages > tc_stage_cap()) L.b_stages = tc_stage_cap();
  L.a_stages = (N <= 64) ? 3 : 2;
  L.tmem_cols = 32;
  while (L.tmem_cols < N + 64 * L.a_stages) L.tmem_cols <<= 1;
  L.raw_off = L.b_stages * L.b_stage_bytes;
  L.pair_off = L.raw_off + 2 * kT16ABytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(2 * chunks, 16);
  L.total = L.bar_off + 512 + 8 * N + 1024;
  return L;
}

template <bool VEC, bool X3>
__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_tc16t_kernel(const float* __restrict__ feat, const uint16_t* __restrict__ wpk,
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
  const int tile = (int)blockIdx.x / split, half = (int)blockIdx.x % split;   // split-K pairs share a tile
  const int row0 = tile * kTcM;
  TC_TRACE_INIT();
  TC_TRACE_ENTRY();
  tc::pdl_launch_dependents();

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&b_full[s], 1);                 // the arrive.expect_tx of the bulk copy
      tc::mbar_init(&b_empty[s], 1);                // tcgen05.commit
      tc::mbar_init(&a_full[s], kTcProducerWarps);  // one arrive per converter warp after its tcgen05.st
      tc::mbar_init(&a_empty[s], 1);                // tcgen05.commit
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
    tc_build_active_list(used_s, chunks, cin_pad, kvol, lane, (unsigned short*)alist, n_act_s, kT16KC);
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
  tc::pdl_wait();

  if (warp < kTcProducerWarps) {
    // ===== gather (coalesced) -> raw smem -> row-per-thread read -> bf16 [hi | lo] -> tcgen05.st =====
    const int p = tid & 7;
    const int rbase = tid >> 3;  // 0..31
    constexpr int RPT = kTcM / (kTcProducers / 8);  // 4 rows per thread in the gather mapping
    auto gather = [&](int t, float4 (&v)[RPT]) {   // t = gather step: chunk alist[t / 2], half t % 2
      const int kk0 = ((int)alist[t >> 1] * 2 + (t & 1)) * kT16Step + p * 4;
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
    // converter mapping: this thread owns accumulator row `crow` (the TMEM lane it may write) and 16 of the
    // step's 32 K elements = 8 of the chunk's 32 packed columns: columns 16*h + 8*chalf .. +7
    const int crow = (warp & 3) * 32 + lane;
    const int chalf = warp >> 2;
    const uint32_t raw0 = tc::smem_u32(smem + raw_off);
    const int tr_role = warp == 0 ? 0 : (warp == kTcProducerWarps - 1 ? 1 : -1);  // traced gather warps
    (void)tr_role;
    auto convert = [&](int t, const float4 (&v)[RPT]) {
      const int it = t >> 1, h = t & 1;
      const uint32_t raw = raw0 + (uint32_t)(t & 1) * kT16ABytes;
      // 1. coalesced-layout registers -> swizzled raw tile (one pass, conflict-free)
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rbase + 32 * i;
        tc::st_shared_v4(raw + (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)), v[i].x, v[i].y, v[i].z, v[i].w);
      }
      tc::named_bar_sync(1, kTcProducers);
      // 2. my row, my 16 K elements (4 pieces), back out of the raw tile -> packed bf16 pairs
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int piece = chalf * 4 + q;
        const float4 x = tc::ld_shared_v4(raw + (uint32_t)(crow * 128 + ((piece ^ (crow & 7)) << 4)));
        const uint32_t h01 = tc::pack_bf16x2(x.x, x.y), h23 = tc::pack_bf16x2(x.z, x.w);
        hi[2 * q] = h01;
        hi[2 * q + 1] = h23;
        if (X3) {
          lo[2 * q] = tc::pack_bf16x2(x.x - __uint_as_float(h01 << 16), x.y - __uint_as_float(h01 & 0xFFFF0000u));
          lo[2 * q + 1] = tc::pack_bf16x2(x.z - __uint_as_float(h23 << 16), x.w - __uint_as_float(h23 & 0xFFFF0000u));
        }
      }
      // 3. TMEM A stage of the chunk free?  (first half-step only)  store, publish after the second half-step
      const int sa = it % a_stages;
      if (h == 0) {
        if (lane == 0) TC_TRACE(tr_role, it, 0);
        mbar_wait_warp(&a_empty[sa], ((uint32_t)(it / a_stages) & 1u) ^ 1u, lane);
        if (lane == 0) TC_TRACE(tr_role, it, 1);
        tc::fence_after_sync();
      }
      const uint32_t ta = tmem_a0 + (uint32_t)(sa * 64) + ((uint32_t)((warp & 3) * 32) << 16) +
                          (uint32_t)(16 * h + 8 * chalf);
      tc::tmem_st8(ta, hi);
      if (X3) tc::tmem_st8(ta + 32u, lo);
      if (h == 1) {
        tc::tmem_st_wait();  // warp-wide: the stores of both half-steps have landed
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&a_full[sa]);
        if (lane == 0) TC_TRACE(tr_role, it, 2);
      }
    };
    const int n_steps = 2 * n_act;
    float4 bufa[RPT], bufb[RPT];
    if (n_steps > 0) gather(0, bufa);
    if (n_steps > 1) gather(1, bufb);
    for (int i = 0; i < n_steps; i += 2) {   // n_steps is even
      convert(i, bufa);
      if (i + 2 < n_steps) gather(i + 2, bufa);
      convert(i + 1, bufb);
      if (i + 3 < n_steps) gather(i + 3, bufb);
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
        tc::bulk_g2s(smem + (size_t)s * b_stage_bytes, (const uint8_t*)wpk + (size_t)j * bytes, bytes, &b_full[s]);
      }
    }
  } else {
    // ===== MMA issuer: A from TMEM (8 columns per K = 16 step), B from shared memory ===============
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtBF16, kTcM, N);
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
        const uint32_t b_lo = b_hi + (uint32_t)N * 128u;
        const uint32_t a_hi = tmem_a0 + (uint32_t)(sa * 64);
        const uint32_t a_lo = a_hi + 32u;
#pragma unroll
        for (int ks = 0; ks < kT16KC / 16; ++ks) {
          const uint64_t dbh = tc::desc_k_sw128(b_hi + (uint32_t)ks * 32u);
          if (X3) {
            const uint64_t dbl = tc::desc_k_sw128(b_lo + (uint32_t)ks * 32u);
            tc::mma_f16_ts(tmem_base, a_lo + (uint32_t)ks * 8u, dbh, idesc, accumulate);
            tc::mma_f16_ts(tmem_base, a_hi + (uint32_t)ks * 8u, dbl, idesc, 1u);
            tc::mma_f16_ts(tmem_base, a_hi + (uint32_t)ks * 8u, dbh, idesc, 1u);
          } else {
            tc::mma_f16_ts(tmem_base, a_hi + (uint32_t)ks * 8u, dbh, idesc, accumulate);
          }
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

// Packed weight image: [chunk j][image: hi (, lo)][n < N][64 bf16, 16-byte units swizzled by (n & 7)].
__global__ void __launch_bounds__(256)
tc16_pack_weight_kernel(const float* __restrict__ w, int cout, int kvol, int cin, int cin_pad, int N,
                        int chunks, int images, uint16_t* __restrict__ packed) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)chunks * N * kT16KC;
  if (t >= total) return;
  const int kk = (int)(t % kT16KC);
  const int n = (int)((t / kT16KC) % N);
  const int j = (int)(t / ((size_t)kT16KC * N));
  const int K = j * kT16KC + kk;
  const int k = K / cin_pad, c = K - k * cin_pad;
  float val = 0.f;
  if (k < kvol && c < cin && n < cout) val = w[((size_t)n * kvol + k) * cin + c];
  const uint32_t hi = tc::pack_bf16x2(val, 0.f) & 0xFFFFu;
  const size_t blk = (size_t)N * kT16KC;
  const size_t off = (size_t)n * kT16KC + (size_t)((((kk >> 3) ^ (n & 7)) << 3) + (kk & 7));
  packed[((size_t)j * images + 0) * blk + off] = (uint16_t)hi;
  if (images == 2) {
    const float lo = val - __uint_as_float(hi << 16);
    packed[((size_t)j * images + 1) * blk + off] = (uint16_t)(tc::pack_bf16x2(lo, 0.f) & 0xFFFFu);
  }
}

struct T16Geom {
  int cin_pad, N, chunks;
};
static bool t16_geom(int cout, int kvol, int cin, T16Geom& g) {
  if (cout < 1 || cout > 256 || kvol < 1 || kvol > 32 || cin < 1) return false;
  g.cin_pad = round_up(cin, 4);
  g.N = round_up(cout, 16);
  g.chunks = ((long long)kvol * g.cin_pad + kT16KC - 1) / kT16KC;
  return true;
}

}  // namespace msmd

using namespace msmd;

#ifdef MSMD_TC_TRACE
extern "C" MSMD_API int msmd_tc16_trace_set(unsigned long long* buf) { return tc_trace_set_impl(buf); }
#endif

extern "C" MSMD_API size_t msmd_spconv_tc16_packed_bytes(int cout, int kvol, int cin, int x3) {
  T16Geom g;
  if (!t16_geom(cout, kvol, cin, g)) return 0;
  return (size_t)g.chunks * (x3 ? 2 : 1) * g.N * kT16KC * sizeof(uint16_t);
}

extern "C" MSMD_API int msmd_spconv_tc16_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                                     int x3, void* packed, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  T16Geom g;
  MSMD_REQUIRE(t16_geom(cout, kvol, cin, g), "spconv_tc16: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(weight_krsc && packed, "spconv_tc16_pack_weight: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed & 15) == 0, "spconv_tc16_pack_weight: packed must be 16-byte aligned");
  const size_t total = (size_t)g.chunks * g.N * kT16KC;
  tc16_pack_weight_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(
      weight_krsc, cout, kvol, cin, g.cin_pad, g.N, g.chunks, x3 ? 2 : 1, (uint16_t*)packed);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

static int g_tc16_variant = 2;  // 2: A through shared memory (default); 3: A through tensor memory (+ split-K pairs)

extern "C" MSMD_API int msmd_spconv_tc16_set_variant(int variant) {
  MSMD_REQUIRE(variant == 2 || variant == 3, "spconv_tc16_set_variant: variant must be 2 or 3");
  g_tc16_variant = variant;
  return MSMD_OK;
}

// split-K scratch of variant 3 (as msmd_spconv_tc_workspace): 0 = no split for this shape / variant
extern "C" MSMD_API size_t msmd_spconv_tc16_workspace(int n_out, int cout) {
  if (g_tc16_variant != 3) return 0;
  const int N = round_up(cout > 0 ? cout : 1, 16);
  const int tiles = ceil_div(n_out > 0 ? n_out : 1, kTcM);
  const bool split = g_tc_tune[2] == 2 ? N >= 96 : (g_tc_tune[2] != 1 && N >= 96 &&
                     (tiles <= kNumSMs / 2 || (tiles > kNumSMs && tiles <= kNumSMs + kNumSMs / 2)));
  if (!split) return 0;
  return (size_t)tiles * kTcM * N * sizeof(float) + (size_t)tiles * 2 * sizeof(int) + 512;
}

static int tc16t_forward(const float* features, const void* packed_tc16, const int* pair_fwd, const int* row_perm,
                         int n_out, int cin, int cout, int kvol, int x3, const T16Geom& g, bool vec,
                         const float* scale, const float* shift, const float* residual, int relu, float* out,
                         void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const int tiles = ceil_div(n_out, kTcM);
  const T16tLayout L = t16t_layout(g.N, kvol, g.chunks, x3 ? 2 : 1);
  MSMD_REQUIRE(L.b_stages >= 1 && L.tmem_cols <= 512, "spconv_fwd_tc16: tile does not fit on the SM");
  auto kern = x3 ? (vec ? spconv_fwd_tc16t_kernel<true, true> : spconv_fwd_tc16t_kernel<false, true>)
                 : (vec ? spconv_fwd_tc16t_kernel<true, false> : spconv_fwd_tc16t_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[2 * (x3 ? 1 : 0) + vec]) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[2 * (x3 ? 1 : 0) + vec] = true;
  }
  int split = 1;
  float* part_ws = nullptr;
  int* part_flag = nullptr;
  const size_t need = msmd_spconv_tc16_workspace(n_out, cout);
  if (workspace && need > 0 && workspace_bytes >= need) {
    split = 2;
    part_flag = (int*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    part_ws = (float*)(part_flag + (size_t)2 * tiles + (64 - (2 * tiles) % 64) % 64);
    MSMD_CUDA_OK(cudaMemsetAsync(part_flag, 0, (size_t)2 * tiles * sizeof(int), stream));
  }
  tc_launch(kern, tiles * split, kTcThreads, L.total, stream, features, (const uint16_t*)packed_tc16, pair_fwd,
            n_out, cin, g.cin_pad, cout, g.N, kvol, g.chunks, L.b_stages, L.b_stage_bytes, L.a_stages, L.raw_off,
            L.pair_off, L.act_off, L.bar_off, L.tmem_cols, scale, shift, residual, relu, out, split, part_ws,
            part_flag, row_perm);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_fwd_tc16_ws(const float* features, int n_in, const void* packed_tc16,
                                                const int* pair_fwd, const int* row_perm, int n_out, int cin,
                                                int cout, int kvol, int x3, const float* scale,
                                                const float* shift, const float* residual, int relu, float* out,
                                                void* workspace, size_t workspace_bytes, msmd_stream_t stream_);

extern "C" MSMD_API int msmd_spconv_fwd_tc16(const float* features, int n_in, const void* packed_tc16,
                                             const int* pair_fwd, const int* row_perm, int n_out, int cin,
                                             int cout, int kvol, int x3, const float* scale,
                                             const float* shift, const float* residual, int relu, float* out,
                                             msmd_stream_t stream_) {
  return msmd_spconv_fwd_tc16_ws(features, n_in, packed_tc16, pair_fwd, row_perm, n_out, cin, cout, kvol, x3, scale,
                                 shift, residual, relu, out, nullptr, 0, stream_);
}

extern "C" MSMD_API int msmd_spconv_fwd_tc16_ws(const float* features, int n_in, const void* packed_tc16,
                                                const int* pair_fwd, const int* row_perm, int n_out, int cin,
                                                int cout, int kvol, int x3, const float* scale,
                                                const float* shift, const float* residual, int relu, float* out,
                                                void* workspace, size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  T16Geom g;
  MSMD_REQUIRE(t16_geom(cout, kvol, cin, g), "spconv_fwd_tc16: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0, "spconv_fwd_tc16: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd_tc16: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(features && packed_tc16 && pair_fwd && out, "spconv_fwd_tc16: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed_tc16 & 15) == 0, "spconv_fwd_tc16: packed weights must be 16-byte aligned");
  const bool vec = (cin % 4 == 0) && (((uintptr_t)features & 15) == 0);
  if (g_tc16_variant == 3)
    return tc16t_forward(features, packed_tc16, pair_fwd, row_perm, n_out, cin, cout, kvol, x3, g, vec, scale, shift,
                         residual, relu, out, workspace, workspace_bytes, stream);
  const int tiles = ceil_div(n_out, kTcM);
  const int images = x3 ? 2 : 1;
  // A/B switch [3]: two chunk blocks per pipeline stage (half the mbarrier round trips) when two such stages fit
  int cps = 1;
  T16Layout L = t16_layout(g.N, kvol, g.chunks, tiles, images, 1);
  if (g_tc_tune[3] == 2) {
    const T16Layout L2 = t16_layout(g.N, kvol, g.chunks, tiles, images, 2);
    if (L2.stages >= 2) { L = L2; cps = 2; }
  }
  MSMD_REQUIRE(L.stages >= 2, "spconv_fwd_tc16: tile does not fit in shared memory");
  auto kern = x3 ? (vec ? spconv_fwd_tc16_kernel<true, true> : spconv_fwd_tc16_kernel<false, true>)
                 : (vec ? spconv_fwd_tc16_kernel<true, false> : spconv_fwd_tc16_kernel<false, false>);
  static bool attr_set[4] = {false, false, false, false};
  if (!attr_set[2 * (x3 ? 1 : 0) + vec]) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[2 * (x3 ? 1 : 0) + vec] = true;
  }
  // concatenated-B mode (x3 only) needs 2N accumulator columns; two co-resident CTAs must fit in 512
  const int cat = (x3 && 2 * g.N <= 256) ? 1 : 0;
  int tmem_cols = 32;
  while (tmem_cols < (cat ? 2 * g.N : g.N)) tmem_cols <<= 1;
  tc_launch(kern, tiles, kTcThreads, L.total, stream, features, (const uint16_t*)packed_tc16, pair_fwd, n_out, cin,
                                               g.cin_pad, cout, g.N, kvol, g.chunks, L.stages, L.stage_bytes, cps,
                                               L.pair_off, L.act_off, L.bar_off, tmem_cols, scale, shift,
                                               residual, relu, out, cat, row_perm);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
