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
// Pipeline: `stages` shared-memory stages, mbarrier full/empty per stage (full = 256 producer
// arrivals + the bulk copy's transaction bytes; empty = tcgen05.commit), one accumulator-ready
// barrier for the epilogue.  Warps 0-7 gather (each thread keeps the loads of TWO chunks in
// flight), then run the epilogue; warp 8 issues the B bulk copies; warp 9 allocates TMEM and
// issues the MMAs.
#include "tc.cuh"

namespace msmd {

constexpr int kTcProducerWarps = 8;                       // gather warps (also the epilogue)
constexpr int kTcProducers = kTcProducerWarps * 32;       // 256 threads
constexpr int kTcThreads = kTcProducers + 64;             // + B-loader warp + MMA warp
constexpr int kTcM = 128;        // output voxels per tile (UMMA M)
constexpr int kTcKC = 32;        // floats per K chunk = one 128-byte swizzle row
constexpr int kTcABytes = kTcM * kTcKC * 4;  // 16 KB per A half (hi or lo)

struct TcSmemLayout {
  int stage_bytes, stages, pair_off, act_off, bar_off, total;
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

static TcSmemLayout tc_layout(int N, int kvol, int chunks) {
  TcSmemLayout L;
  L.stage_bytes = 2 * kTcABytes + N * kTcKC * 4 * 2;
  const int misc = round_up(kvol * kTcM * 4, 16) + round_up(chunks, 16) + 256;
  // two CTAs per SM when two stages fit in half of the SM's shared memory, else one CTA
  // with as many stages as fit (max 4)
  const int half = 112 * 1024, full = 224 * 1024;
  if (2 * L.stage_bytes + misc + 1024 <= half) {
    L.stages = (half - misc - 1024) / L.stage_bytes;
  } else {
    L.stages = (full - misc - 1024) / L.stage_bytes;
  }
  if (L.stages > 4) L.stages = 4;
  L.pair_off = L.stages * L.stage_bytes;
  L.act_off = L.pair_off + round_up(kvol * kTcM * 4, 16);
  L.bar_off = L.act_off + round_up(chunks, 16);
  L.total = L.bar_off + 256 + 1024;  // + slack for the 1024-byte alignment of the base
  return L;
}

template <bool VEC>
__global__ void __launch_bounds__(kTcThreads)
spconv_fwd_tc_kernel(const float* __restrict__ feat, const float* __restrict__ wpk,
                     const int* __restrict__ pair, int n_out, int cin, int cin_pad, int cout, int N,
                     int kvol, int chunks, int stages, int stage_bytes, int pair_off, int act_off,
                     int bar_off, int tmem_cols, const float* __restrict__ scale,
                     const float* __restrict__ shift, const float* __restrict__ residual, int relu,
                     float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  int* pair_s = (int*)(smem + pair_off);
  uint8_t* act = smem + act_off;
  uint64_t* full_bar = (uint64_t*)(smem + bar_off);
  uint64_t* empty_bar = full_bar + 4;
  uint64_t* accum_bar = full_bar + 8;
  uint32_t* tmem_ptr_s = (uint32_t*)(full_bar + 9);
  int* used_s = (int*)(full_bar + 10);  // [kvol <= 32]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTcM;

  // ---- one-time setup -------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], kTcProducers + 1);  // gather threads + the arrive.expect_tx of the B copy
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
  int mine = 0;
  for (int j = tid; j < chunks; j += kTcThreads) {
    const int k_lo = (j * kTcKC) / cin_pad;
    int k_hi = (j * kTcKC + kTcKC - 1) / cin_pad;
    if (k_hi > kvol - 1) k_hi = kvol - 1;
    int a = 0;
    for (int k = k_lo; k <= k_hi; ++k) a |= used_s[k];
    act[j] = (uint8_t)a;
    mine |= a;
  }
  const int any_active = __syncthreads_or(mine);
  const uint32_t tmem_base = *tmem_ptr_s;

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
    auto next_active = [&](int j) {
      ++j;
      while (j < chunks && !act[j]) ++j;
      return j;
    };
    int it = 0;
    auto store = [&](const float4 (&v)[RPT]) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)(it / stages) & 1u;
      tc::mbar_wait(&empty_bar[s], ph ^ 1u);
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
      tc::fence_proxy_async();
      tc::mbar_arrive(&full_bar[s]);
      ++it;
    };
    // two register buffers, alternated without copies: while one chunk is split and stored the
    // loads of the following chunk are already in flight
    float4 bufa[RPT], bufb[RPT];
    int ja = next_active(-1);
    int jb = chunks;
    if (ja < chunks) {
      gather(ja, bufa);
      jb = next_active(ja);
      if (jb < chunks) gather(jb, bufb);
    }
    while (ja < chunks) {
      store(bufa);
      ja = (jb < chunks) ? next_active(jb) : chunks;
      if (ja < chunks) gather(ja, bufa);
      if (jb >= chunks) break;
      store(bufb);
      jb = (ja < chunks) ? next_active(ja) : chunks;
      if (jb < chunks) gather(jb, bufb);
    }

    // ===== epilogue: TMEM -> registers -> fused BN / residual / ReLU -> global =============
    if (any_active) {
      tc::mbar_wait(accum_bar, 0);
      tc::fence_after_sync();
    }
    const int quarter = warp & 3;             // TMEM lanes this warp may read: 32*quarter ..
    const int o = row0 + quarter * 32 + lane;
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
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0u;
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
            if (scale && co + q < cout) y[q] = fmaf(y[q], __ldg(scale + co + q), __ldg(shift + co + q));
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
  } else if (warp == kTcProducerWarps) {
    // ===== B loader: one bulk copy (hi + lo image of the chunk) per active chunk ============
    if (lane == 0) {
      const uint32_t bytes = (uint32_t)N * kTcKC * 4u * 2u;
      int it = 0;
      for (int j = 0; j < chunks; ++j) {
        if (!act[j]) continue;
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        tc::mbar_wait(&empty_bar[s], ph ^ 1u);
        tc::mbar_arrive_expect_tx(&full_bar[s], bytes);
        tc::bulk_g2s(smem + (size_t)s * stage_bytes + 2 * kTcABytes,
                     wpk + (size_t)j * N * kTcKC * 2, bytes, &full_bar[s]);
        ++it;
      }
    }
  } else {
    // ===== MMA issuer ========================================================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_f32acc(tc::kFmtTF32, kTcM, N);
      uint32_t accumulate = 0;
      int it = 0;
      for (int j = 0; j < chunks; ++j) {
        if (!act[j]) continue;
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        tc::mbar_wait(&full_bar[s], ph);
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
          tc::mma_tf32(tmem_base, dal, dbh, idesc, accumulate);  // small terms first
          tc::mma_tf32(tmem_base, dah, dbl, idesc, 1u);
          tc::mma_tf32(tmem_base, dah, dbh, idesc, 1u);
          accumulate = 1u;
        }
        tc::mma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
        ++it;
      }
      if (it > 0) tc::mma_commit(accum_bar);  // accumulator complete -> epilogue
    }
  }

  // ---- teardown -------------------------------------------------------------------------
  tc::fence_before_sync();
  __syncthreads();
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
  uint32_t hb;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(val));
  const float hi = __uint_as_float(hb);
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

extern "C" MSMD_API int msmd_spconv_fwd_tc(const float* features, int n_in, const float* packed_tc,
                                           const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                           const float* scale, const float* shift,
                                           const float* residual, int relu, float* out,
                                           msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TcGeom g;
  MSMD_REQUIRE(tc_geom(cout, kvol, cin, g), "spconv_fwd_tc: unsupported shape (cout<=256, kvol<=32)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0, "spconv_fwd_tc: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd_tc: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(features && packed_tc && pair_fwd && out, "spconv_fwd_tc: null pointer");
  MSMD_REQUIRE(((uintptr_t)packed_tc & 15) == 0, "spconv_fwd_tc: packed weights must be 16-byte aligned");
  const TcSmemLayout L = tc_layout(g.N, kvol, g.chunks);
  MSMD_REQUIRE(L.stages >= 2, "spconv_fwd_tc: tile does not fit in shared memory");
  const bool vec = (cin % 4 == 0) && (((uintptr_t)features & 15) == 0);
  const int tiles = ceil_div(n_out, kTcM);
  static bool attr_set[2] = {false, false};
  auto kern = vec ? spconv_fwd_tc_kernel<true> : spconv_fwd_tc_kernel<false>;
  if (!attr_set[vec]) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[vec] = true;
  }
  kern<<<tiles, kTcThreads, L.total, stream>>>(features, packed_tc, pair_fwd, n_out, cin, g.cin_pad, cout,
                                               g.N, kvol, g.chunks, L.stages, L.stage_bytes, L.pair_off,
                                               L.act_off, L.bar_off, g.tmem_cols, scale, shift, residual,
                                               relu, out);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
