// spconv_wgrad_tc.cu -- weight gradient of the sparse convolution on the tensor cores (tcgen05, 3xTF32).
//
//     dW[co, k, ci] = sum over the pairs (i -> o) of kernel offset k of  dY[o, co] * X[i, ci]
//
// (spconv-1.x indiceConvBackward, mmdet3d/ops/spconv/include/spconv/spconv_ops.h:364-457; the reference
// reaches it through the backward of Fsp.implicit_gemm, bug_fix/conv.py:442-447.)  Per offset this is a
// GEMM  D[M = co, N = ci] = A[M, K] * B[N, K]^T  whose reduction axis K runs over the offset's PAIRS:
// A[co, p] = dY[o_p, co], B[ci, p] = X[i_p, ci] -- both operands are the gathered rows TRANSPOSED.
//
// One CTA owns (row slice, kernel offset k, 128-row co tile) like the SIMT kernel it replaces
// (spconv_bwd.cu): it scans its slice of pair_fwd[k] 1024 rows at a time, compacts the active pairs
// (block scan, ascending row order), and feeds them to the tensor core 32 pairs (one K chunk) at a time:
//   1. gather: coalesced float4 loads of dY[o_p, co tile] and X[i_p, :] into registers (issued one chunk
//      ahead), stored row-major into "raw" shared-memory tiles [32 pairs][channels];
//   2. transpose + split: thread = channel reads its COLUMN of the raw tile (conflict-free: consecutive
//      lanes, consecutive words), splits the 32 values into tf32 hi / lo and writes
//        A (co rows)  -> tensor memory with tcgen05.st (lane = co, 32 + 32 columns),
//        B (ci rows)  -> shared memory, one 128-byte K-major SWIZZLE_128B row per ci (hi image, lo image);
//   3. one thread issues 4 k-steps x 3 MMAs (A_lo*B_hi, A_hi*B_lo, A_hi*B_hi) into the fp32 accumulator
//      D[128, N] in tensor memory and commits to an mbarrier that step 2 of the NEXT chunk waits on.
// The accumulator lives in TMEM across the whole slice; at the end it is written as a partial tile to the
// workspace [slice][k][co][ci] and a second kernel sums the slices in a fixed order: deterministic, no
// atomics.  fp32-level accuracy (3xTF32, lo*lo dropped) instead of the SIMT kernel's exact FFMA.
//
// Status: written without GPU time; opt-in (msmd_spconv_set_wgrad_tc / MSMD_WGRAD_TC=1).  Correct on a B200
// (tools/quick_gpu_check.py, profiles/r01h_quick_gpu_check.json: <= 8e-6 of float64), not yet timed.  Checked on the
// host model of tcgen05 (tests/test_cuda_emul.py::test_wgrad_tc_*); GPU test
// tests/test_zz_train_gpu.py::test_wgrad_tc_matches_simt_and_oracle.
#include "tc.cuh"

namespace msmd {

constexpr int kWtThreads = 256;
constexpr int kWtM = 128;           // co rows per CTA tile (UMMA M)
constexpr int kWtKC = 32;           // pairs per K chunk (one 128-byte row of tf32)
constexpr int kWtScan = 1024;       // output rows scanned per compaction step (4 per thread)
constexpr int kWtMaxN = 256;        // ci extent (UMMA N) a CTA covers: all of cin

struct WtLayout {
  int raw_a_off, raw_b_off, b_hi_off, b_lo_off, list_off, misc_off, total;
};
static WtLayout wt_layout(int N) {
  WtLayout L;
  L.b_hi_off = 0;                         // N rows x 128 B, 1024-B aligned
  L.b_lo_off = L.b_hi_off + N * 128;
  L.raw_a_off = L.b_lo_off + N * 128;     // [32][128] floats
  L.raw_b_off = L.raw_a_off + kWtKC * kWtM * 4;   // [32][N] floats
  L.list_off = L.raw_b_off + kWtKC * N * 4;       // list_o[1024] | list_i[1024]
  L.misc_off = L.list_off + 2 * kWtScan * 4;      // mbarrier | tmem ptr | scan scratch (33 ints)
  L.total = L.misc_off + 256 + 1024;              // + slack for the 1024-byte alignment
  return L;
}

__global__ void __launch_bounds__(kWtThreads)
spconv_wgrad_tc_kernel(const float* __restrict__ feat, const float* __restrict__ grad_out,
                       const int* __restrict__ pair, int n_out, int cin, int cout, int kvol, int N,
                       int n_split, int tmem_cols, int raw_a_off, int raw_b_off, int b_lo_off, int list_off,
                       int misc_off, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* raw_a = (float*)(smem + raw_a_off);     // [32][128]: dY rows of the chunk's pairs (co tile)
  float* raw_b = (float*)(smem + raw_b_off);     // [32][N]:   X rows of the chunk's pairs
  int* list_o = (int*)(smem + list_off);
  int* list_i = list_o + kWtScan;
  uint64_t* mma_bar = (uint64_t*)(smem + misc_off);
  uint32_t* tmem_ptr_s = (uint32_t*)(mma_bar + 1);
  int* scan_s = (int*)(mma_bar + 2);             // 33 ints

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x;
  const int k = blockIdx.y;
  const int co0 = (int)blockIdx.z * kWtM;

  if (tid == 0) {
    tc::mbar_init(mma_bar, 1);   // one tcgen05.commit per K chunk
    tc::fence_mbar_init();
  }
  if (warp == 0) {
    tc::tmem_alloc(tmem_ptr_s, (uint32_t)tmem_cols);
    tc::tmem_relinquish();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_d = *tmem_ptr_s;            // accumulator: columns [0, N)
  const uint32_t tmem_a = tmem_d + (uint32_t)N;   // A operand: hi 32 columns | lo 32 columns
  const uint32_t b_hi = tc::smem_u32(smem);
  const uint32_t b_lo = b_hi + (uint32_t)b_lo_off;
  const uint32_t idesc = tc::idesc_f32acc(tc::kFmtTF32, kWtM, N);

  // gather mapping: raw_a = 32 rows x 32 float4 (4 per thread), raw_b = 32 rows x N/4 float4 (<= 8 per thread)
  const int nb4 = N >> 2;                      // float4 per raw_b row
  constexpr int kA4 = kWtKC * (kWtM / 4) / kWtThreads;   // 4
  constexpr int kB4 = kWtKC * (kWtMaxN / 4) / kWtThreads;  // 8
  float4 ra[kA4], rb[kB4];
  auto gather = [&](int r0, int cnt) {  // rows r0 .. r0+31 of the compacted list (zeros past cnt / past the channels)
#pragma unroll
    for (int q = 0; q < kA4; ++q) {
      const int e = tid + q * kWtThreads, r = e >> 5, c4 = e & 31;
      const int co = co0 + c4 * 4;
      ra[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < cnt && co < cout) ra[q] = __ldg((const float4*)(grad_out + (size_t)list_o[r0 + r] * cout + co));
    }
#pragma unroll
    for (int q = 0; q < kB4; ++q) {
      const int e = tid + q * kWtThreads;
      rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < kWtKC * nb4) {
        const int r = e / nb4, c4 = e - r * nb4;
        if (r0 + r < cnt && c4 * 4 < cin) rb[q] = __ldg((const float4*)(feat + (size_t)list_i[r0 + r] * cin + c4 * 4));
      }
    }
  };
  auto stage = [&]() {  // registers -> raw tiles (row-major, every element defined)
#pragma unroll
    for (int q = 0; q < kA4; ++q) {
      const int e = tid + q * kWtThreads;
      *(float4*)(raw_a + (size_t)e * 4) = ra[q];
    }
#pragma unroll
    for (int q = 0; q < kB4; ++q) {
      const int e = tid + q * kWtThreads;
      if (e < kWtKC * nb4) *(float4*)(raw_b + (size_t)e * 4) = rb[q];
    }
  };

  uint32_t n_mma = 0;          // K chunks issued so far (mma_bar completes once per chunk)
  const int* pk = pair + (size_t)k * n_out;
  const int n_scans = (n_out + kWtScan - 1) / kWtScan;
  for (int c = split; c < n_scans; c += n_split) {   // interleaved slices: balanced pair density
    // ---- compact the active pairs of rows [c*1024, +1024): thread t owns rows 4t .. 4t+3 ----
    int pv[4], valid = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = c * kWtScan + tid * 4 + q;
      pv[q] = (o < n_out) ? __ldg(pk + o) : -1;
      valid += pv[q] >= 0;
    }
    int cnt;
    int pos = block_exclusive_scan<int>(valid, cnt, scan_s);  // leading barrier inside
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (pv[q] >= 0) {
        list_o[pos] = c * kWtScan + tid * 4 + q;
        list_i[pos] = pv[q];
        ++pos;
      }
    __syncthreads();
    if (cnt > 0) gather(0, cnt);
    for (int r0 = 0; r0 < cnt; r0 += kWtKC) {
      stage();                                        // chunk r0: registers -> raw tiles
      if (r0 + kWtKC < cnt) gather(r0 + kWtKC, cnt);  // next chunk's loads fly during the transposition
      __syncthreads();
      // the previous chunk's MMAs have read TMEM A and the B images?
      if (n_mma > 0) {
        if (lane == 0) tc::mbar_wait(mma_bar, (n_mma - 1) & 1u);
        __syncwarp();
        tc::fence_after_sync();
      }
      // ---- transpose + split ----
      if (warp < 4) {   // A: lane of TMEM = co row (warp w may touch lanes 32w ..)
        const int m = tid;   // 0..127
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x = raw_a[(half * 16 + j) * kWtM + m];
            const float h = tc::round_tf32(x);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(x - h);
          }
          const uint32_t ta = tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)(half * 16);
          tc::tmem_st16(ta, hi);
          tc::tmem_st16(ta + 32u, lo);
        }
        tc::tmem_st_wait();
      }
      for (int n = tid; n < N; n += kWtThreads) {   // B: one 128-byte swizzled row per ci
#pragma unroll
        for (int u = 0; u < 8; ++u) {   // 16-byte unit u = pairs 4u .. 4u+3
          float x[4], h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            x[j] = raw_b[(u * 4 + j) * N + n];
            h[j] = tc::round_tf32(x[j]);
          }
          const uint32_t off = (uint32_t)(n * 128 + ((u ^ (n & 7)) << 4));
          tc::st_shared_v4(b_hi + off, h[0], h[1], h[2], h[3]);
          tc::st_shared_v4(b_lo + off, x[0] - h[0], x[1] - h[1], x[2] - h[2], x[3] - h[3]);
        }
      }
      tc::fence_proxy_async();     // generic-proxy stores of B -> async proxy
      tc::fence_before_sync();     // tcgen05.st of A ordered before the barrier
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
#pragma unroll
        for (int ks = 0; ks < kWtKC / 8; ++ks) {
          const uint64_t dbh = tc::desc_k_sw128(b_hi + (uint32_t)ks * 32u);
          const uint64_t dbl = tc::desc_k_sw128(b_lo + (uint32_t)ks * 32u);
          const uint32_t a_hi = tmem_a + (uint32_t)ks * 8u, a_lo = a_hi + 32u;
          tc::mma_tf32_ts(tmem_d, a_lo, dbh, idesc, (n_mma | (uint32_t)ks) ? 1u : 0u);  // small terms first
          tc::mma_tf32_ts(tmem_d, a_hi, dbl, idesc, 1u);
          tc::mma_tf32_ts(tmem_d, a_hi, dbh, idesc, 1u);
        }
        tc::mma_commit(mma_bar);
      }
      ++n_mma;
    }
  }

  // ---- accumulator -> partial tile [split][k][co][ci] (zeros when the slice had no pair) ----
  if (n_mma > 0) {
    if (lane == 0) tc::mbar_wait(mma_bar, (n_mma - 1) & 1u);
    __syncwarp();
    tc::fence_after_sync();
  }
  float* dst = partial + ((size_t)split * kvol + k) * cout * cin;
  const int quarter = warp & 3;
  const int co = co0 + quarter * 32 + lane;
  const int nsteps = N / 16;
  const int step_lo = (warp >> 2) ? (nsteps + 1) / 2 : 0;   // two warpgroups split the columns
  const int step_hi = (warp >> 2) ? nsteps : (nsteps + 1) / 2;
  for (int st = step_lo; st < step_hi; ++st) {
    uint32_t acc[16];
    if (n_mma > 0) {
      tc::tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(st * 16), acc);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[e] = 0u;
    }
    if (co < cout) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int ci = st * 16 + e;
        if (ci < cin) dst[(size_t)co * cin + ci] = __uint_as_float(acc[e]);
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

__global__ void __launch_bounds__(256)
spconv_wgrad_tc_reduce_kernel(const float* __restrict__ partial, int n_split, int cout, int kvol, int cin,
                              float* __restrict__ grad_w) {
  // t enumerates the KRSC gradient [co][k][ci]; slices are summed in a fixed order
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)cout * kvol * cin;
  if (t >= total) return;
  const int ci = (int)(t % cin);
  const int k = (int)((t / cin) % kvol);
  const int co = (int)(t / ((size_t)cin * kvol));
  const size_t src = ((size_t)k * cout + co) * cin + ci;
  float s = 0.f;
  for (int sp = 0; sp < n_split; ++sp) s += partial[(size_t)sp * total + src];
  grad_w[t] = s;
}

static int wgrad_tc_splits(int n_out, int cout, int kvol) {
  const long long tiles = (long long)kvol * ceil_div(cout, kWtM);
  const int scans = ceil_div(n_out, kWtScan);
  long long s = (2LL * kNumSMs + tiles - 1) / tiles;  // ~2 CTAs per SM in flight
  if (s > scans) s = scans;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  return (int)s;
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API int msmd_spconv_bwd_weight_tc_supported(int cin, int cout, int kvol) {
  return (cin >= 4 && cin <= kWtMaxN && cin % 4 == 0 && cout >= 4 && cout % 4 == 0 && kvol >= 1) ? 1 : 0;
}

extern "C" MSMD_API size_t msmd_spconv_bwd_weight_tc_workspace(int n_out, int cin, int cout, int kvol) {
  if (n_out <= 0 || !msmd_spconv_bwd_weight_tc_supported(cin, cout, kvol)) return 0;
  return (size_t)wgrad_tc_splits(n_out, cout, kvol) * kvol * cin * cout * sizeof(float);
}

extern "C" MSMD_API int msmd_spconv_bwd_weight_tc(const float* features, int n_in, const float* grad_out,
                                                  const int* pair_fwd, int n_out, int cin, int cout,
                                                  int kvol, float* grad_weight_krsc, void* workspace,
                                                  size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(msmd_spconv_bwd_weight_tc_supported(cin, cout, kvol),
               "spconv_bwd_weight_tc: unsupported shape (cin, cout multiples of 4, cin <= 256)");
  MSMD_REQUIRE(n_in >= 0 && n_out >= 0 && grad_weight_krsc, "spconv_bwd_weight_tc: bad arguments");
  const size_t total = (size_t)cout * kvol * cin;
  if (n_out == 0 || n_in == 0) {
    MSMD_CUDA_OK(cudaMemsetAsync(grad_weight_krsc, 0, total * sizeof(float), stream));
    return MSMD_OK;
  }
  MSMD_REQUIRE(features && grad_out && pair_fwd, "spconv_bwd_weight_tc: null pointer");
  MSMD_REQUIRE((((uintptr_t)features | (uintptr_t)grad_out) & 15) == 0,
               "spconv_bwd_weight_tc: features / grad_out must be 16-byte aligned");
  const int splits = wgrad_tc_splits(n_out, cout, kvol);
  if (workspace == nullptr || workspace_bytes < (size_t)splits * total * sizeof(float)) {
    set_error("spconv_bwd_weight_tc: workspace too small (%zu < %zu bytes)", workspace_bytes,
              (size_t)splits * total * sizeof(float));
    return MSMD_ERR_WORKSPACE;
  }
  const int N = (cin + 15) / 16 * 16;
  const WtLayout L = wt_layout(N);
  int tmem_cols = 32;
  while (tmem_cols < N + 64) tmem_cols <<= 1;
  static bool attr_set = false;
  if (!attr_set) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(spconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024));
    attr_set = true;
  }
  float* partial = (float*)workspace;
  dim3 grid(splits, kvol, ceil_div(cout, kWtM));
  spconv_wgrad_tc_kernel<<<grid, kWtThreads, L.total, stream>>>(
      features, grad_out, pair_fwd, n_out, cin, cout, kvol, N, splits, tmem_cols, L.raw_a_off, L.raw_b_off,
      L.b_lo_off, L.list_off, L.misc_off, partial);
  MSMD_LAUNCH_OK();
  spconv_wgrad_tc_reduce_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(
      partial, splits, cout, kvol, cin, grad_weight_krsc);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
