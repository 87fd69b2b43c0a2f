// spconv_bwd.cu -- backward of the sparse convolution (config 5 of BASELINE.json: the train step).
//
// The reference's train step differentiates spconv-2.x's implicit GEMM through pair_bwd /
// mask_argsort_bwd_splits (Fsp.implicit_gemm backward, call site bug_fix/conv.py:442-447); the
// arithmetic is the one the vendored spconv-1.x spells out (indiceConvBackward,
// mmdet3d/ops/spconv/include/spconv/spconv_ops.h:364-457): per kernel offset k, over the offset's
// pairs (input row i -> output row o)
//
//     dX[i, ci]      += sum_co dY[o, co] * W[co, k, ci]          (dgrad)
//     dW[co, k, ci]   = sum_{pairs of k} dY[o, co] * X[i, ci]     (wgrad)
//
// dgrad.  With pair_bwd[k, i] = o (the transposed rulebook; at most one o per (k, i)) dgrad IS the
// forward contraction:  dX[i, :] = sum_k dY[pair_bwd[k, i], :] . Wt[:, k, :],  Wt[ci, k, co] = W[co, k, ci]
// -- so msmd_spconv_bwd_data runs the forward kernels (tcgen05 path included) on dY with the
// transposed weight, and only two small helpers are new: the rulebook transposition and the weight
// transposition.  For SubM layers pair_bwd[k] = pair_fwd[K-1-k] (o reads i through offset k <=> i
// reads o through the mirrored offset), so the host passes pair_fwd itself together with a weight
// that is transposed AND reversed along k; no second table is built.
//
// wgrad.  One CTA owns a 64 (co) x 64 (ci) tile of one kernel offset's dW_k and a slice of the
// output rows; it compacts the slice's active pairs of that offset (block scan), stages 16-row slabs
// of dY[o] and X[i] in shared memory (float4 loads of whole row segments) and accumulates the outer
// products in registers (4 x 4 per thread, exact fp32 FFMA).  Row slices write partial tiles to a
// workspace that a second kernel sums in a fixed order into the KRSC gradient -- no atomics, so the
// result is deterministic.  Algorithmic bytes:  4*(P*(cin + cout)) gathered + 4*K*N_out pair reads
// + 4*K*cin*cout written; flops 2*P*cin*cout (P = active pairs).
#include "common.cuh"

namespace msmd {

// ---------------------------------------------------------------------------------------------
// rulebook / weight transposition
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pair_transpose_kernel(const int* __restrict__ pair_fwd, int kvol, int n_out, int n_in,
                      int* __restrict__ pair_bwd) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)kvol * n_out) return;
  const int p = __ldg(pair_fwd + t);
  if (p < 0 || p >= n_in) return;
  const int k = (int)(t / n_out), o = (int)(t % n_out);
  pair_bwd[(size_t)k * n_in + p] = o;  // (k, i) is written by at most one (k, o)
}

__global__ void __launch_bounds__(256)
weight_transpose_kernel(const float* __restrict__ w, int cout, int kvol, int cin, int flip_k,
                        float* __restrict__ wt) {
  // t enumerates the OUTPUT layout [ci][k'][co] so the stores are coalesced
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)cout * kvol * cin) return;
  const int co = (int)(t % cout);
  const int kk = (int)((t / cout) % kvol);
  const int ci = (int)(t / ((size_t)cout * kvol));
  const int k = flip_k ? kvol - 1 - kk : kk;
  wt[t] = __ldg(w + ((size_t)co * kvol + k) * cin + ci);
}

// ---------------------------------------------------------------------------------------------
// wgrad
// ---------------------------------------------------------------------------------------------
constexpr int kWgThreads = 256;
constexpr int kWgTile = 64;    // co and ci extent of a CTA tile
constexpr int kWgSlab = 16;    // pair rows staged per step
constexpr int kWgChunk = 256;  // output rows scanned per compaction step (= threads)

template <bool VEC4>
__global__ void __launch_bounds__(kWgThreads)
spconv_wgrad_simt_kernel(const float* __restrict__ feat, const float* __restrict__ grad_out,
                         const int* __restrict__ pair, int n_out, int cin, int cout, int kvol,
                         int tiles_ci, int n_split, float* __restrict__ partial) {
  __shared__ __align__(16) float DY_s[kWgSlab][kWgTile];
  __shared__ __align__(16) float X_s[kWgSlab][kWgTile];
  __shared__ int list_o[kWgChunk];
  __shared__ int list_i[kWgChunk];
  __shared__ int scan_s[33];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // ci quad, co quad
  const int split = blockIdx.x;
  const int k = blockIdx.y;
  const int co0 = ((int)blockIdx.z / tiles_ci) * kWgTile;
  const int ci0 = ((int)blockIdx.z % tiles_ci) * kWgTile;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int* pk = pair + (size_t)k * n_out;
  const int n_chunks = (n_out + kWgChunk - 1) / kWgChunk;
  for (int c = split; c < n_chunks; c += n_split) {  // interleaved slices: balanced pair density
    const int o = c * kWgChunk + tid;
    const int p = (o < n_out) ? __ldg(pk + o) : -1;
    const int valid = p >= 0;
    int cnt;
    const int pos = block_exclusive_scan<int>(valid, cnt, scan_s);  // leading barrier inside
    if (valid) {
      list_o[pos] = o;
      list_i[pos] = p;
    }
    __syncthreads();
    for (int r0 = 0; r0 < cnt; r0 += kWgSlab) {
      // ---- stage kWgSlab rows of dY[:, co0:co0+64] and X[:, ci0:ci0+64] ----
      if (VEC4) {
        const int r = tid >> 4, q = tid & 15;  // 16 rows x 16 float4
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (r0 + r < cnt) {
          const int cc = co0 + q * 4, ic = ci0 + q * 4;
          if (cc < cout) a = __ldg((const float4*)(grad_out + (size_t)list_o[r0 + r] * cout + cc));
          if (ic < cin) b = __ldg((const float4*)(feat + (size_t)list_i[r0 + r] * cin + ic));
        }
        *(float4*)&DY_s[r][q * 4] = a;
        *(float4*)&X_s[r][q * 4] = b;
      } else {
        for (int s = tid; s < kWgSlab * kWgTile; s += kWgThreads) {
          const int r = s / kWgTile, j = s % kWgTile;
          float a = 0.f, b = 0.f;
          if (r0 + r < cnt) {
            if (co0 + j < cout) a = __ldg(grad_out + (size_t)list_o[r0 + r] * cout + co0 + j);
            if (ci0 + j < cin) b = __ldg(feat + (size_t)list_i[r0 + r] * cin + ci0 + j);
          }
          DY_s[r][j] = a;
          X_s[r][j] = b;
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kWgSlab; ++r) {
        const float4 a = *(const float4*)&DY_s[r][ty * 4];
        const float4 b = *(const float4*)&X_s[r][tx * 4];
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- partial tile -> workspace [split][k][co][ci] ----
  float* dst = partial + ((size_t)split * kvol + k) * cout * cin;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < cin) dst[(size_t)co * cin + ci] = acc[i][j];
    }
  }
}

__global__ void __launch_bounds__(256)
spconv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_split, int cout, int kvol, int cin,
                           float* __restrict__ grad_w) {
  // t enumerates the KRSC gradient [co][k][ci]
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)cout * kvol * cin;
  if (t >= total) return;
  const int ci = (int)(t % cin);
  const int k = (int)((t / cin) % kvol);
  const int co = (int)(t / ((size_t)cin * kvol));
  const size_t src = ((size_t)k * cout + co) * cin + ci;
  float s = 0.f;
  for (int sp = 0; sp < n_split; ++sp) s += partial[(size_t)sp * total + src];  // fixed order
  grad_w[t] = s;
}

static int wgrad_splits(int n_out, int cin, int cout, int kvol) {
  const long long tiles = (long long)kvol * ceil_div(cout, kWgTile) * ceil_div(cin, kWgTile);
  const int chunks = ceil_div(n_out, kWgChunk);
  long long s = (4LL * kNumSMs + tiles - 1) / tiles;  // ~4 CTAs per SM in flight
  if (s > chunks) s = chunks;
  if (s > 32) s = 32;
  if (s < 1) s = 1;
  return (int)s;
}

// ---------------------------------------------------------------------------------------------
// dense() backward: gather the active rows out of a (B, C, D, H, W) gradient
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
from_dense_kernel(const int4* __restrict__ indices, const float* __restrict__ dense, int n, int C,
                  int batch, int D, int H, int W, float* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * C) return;
  const int i = (int)(t % n), c = (int)(t / n);  // consecutive threads = consecutive voxels (as to_dense)
  const int4 p = indices[i];
  float v = 0.f;
  if ((unsigned)p.x < (unsigned)batch && (unsigned)p.y < (unsigned)D && (unsigned)p.z < (unsigned)H &&
      (unsigned)p.w < (unsigned)W)
    v = __ldg(dense + ((((size_t)p.x * C + c) * D + p.y) * H + p.z) * W + p.w);
  out[(size_t)i * C + c] = v;
}

// rank of each voxel in a bit grid (= its row in the grid's ascending order); -1 when absent
__global__ void __launch_bounds__(256)
grid_rows_kernel(const int4* __restrict__ indices, int n, int batch, int D, int H, int W,
                 const uint32_t* __restrict__ bits, const int* __restrict__ prefix,
                 int* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = indices[i];
  int r = -1;
  if ((unsigned)c.x < (unsigned)batch && (unsigned)c.y < (unsigned)D && (unsigned)c.z < (unsigned)H &&
      (unsigned)c.w < (unsigned)W) {
    const int L = ((c.x * D + c.y) * H + c.z) * W + c.w;
    const uint32_t word = __ldg(bits + (L >> 5));
    const unsigned b = (unsigned)L & 31u;
    if ((word >> b) & 1u) r = __ldg(prefix + (L >> 5)) + __popc(word & ((1u << b) - 1u));
  }
  rows[i] = r;
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API int msmd_rulebook_transpose(const int* pair_fwd, int kvol, int n_out, int n_in,
                                                int* pair_bwd, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(kvol > 0 && n_out >= 0 && n_in >= 0, "rulebook_transpose: bad sizes");
  if (n_in == 0) return MSMD_OK;
  MSMD_REQUIRE(pair_bwd && (pair_fwd || n_out == 0), "rulebook_transpose: null pointer");
  MSMD_CUDA_OK(cudaMemsetAsync(pair_bwd, 0xff, (size_t)kvol * n_in * sizeof(int), stream));  // -1
  if (n_out == 0) return MSMD_OK;
  pair_transpose_kernel<<<ceil_div((long long)kvol * n_out, 256), 256, 0, stream>>>(
      pair_fwd, kvol, n_out, n_in, pair_bwd);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_transpose_weight(const float* weight_krsc, int cout, int kvol,
                                                     int cin, int flip_k, float* weight_t,
                                                     msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(weight_krsc && weight_t && cout > 0 && kvol > 0 && cin > 0,
               "transpose_weight: bad args");
  weight_transpose_kernel<<<ceil_div((long long)cout * kvol * cin, 256), 256, 0, stream>>>(
      weight_krsc, cout, kvol, cin, flip_k ? 1 : 0, weight_t);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_bwd_data(const float* grad_out, int n_out, const float* packed_wt,
                                             int weight_tc, const int* pair_bwd, int n_in, int cin,
                                             int cout, int kvol, float* grad_in, void* workspace,
                                             size_t workspace_bytes, msmd_stream_t stream) {
  // forward contraction with the channel roles swapped: "cin" = cout, "cout" = cin
  if (weight_tc == 2 || weight_tc == 3)  // bf16x3 / bf16 image (csrc/spconv_tc16.cu)
    return msmd_spconv_fwd_tc16(grad_out, n_out, packed_wt, pair_bwd, nullptr, n_in, cout, cin, kvol,
                                weight_tc == 2, nullptr, nullptr, nullptr, 0, grad_in, stream);
  if (weight_tc)
    return msmd_spconv_fwd_tc_ws(grad_out, n_out, packed_wt, pair_bwd, n_in, cout, cin, kvol, nullptr,
                                 nullptr, nullptr, 0, grad_in, workspace, workspace_bytes, stream);
  return msmd_spconv_fwd(grad_out, n_out, packed_wt, pair_bwd, n_in, cout, cin, kvol, nullptr, nullptr,
                         nullptr, 0, grad_in, stream);
}

static int g_wgrad_tc = 0;  // msmd_spconv_set_wgrad_tc: route msmd_spconv_bwd_weight to the tensor-core kernel

extern "C" MSMD_API int msmd_spconv_set_wgrad_tc(int enable) {
  g_wgrad_tc = enable ? 1 : 0;
  return MSMD_OK;
}

extern "C" MSMD_API size_t msmd_spconv_bwd_weight_workspace(int n_out, int cin, int cout, int kvol) {
  if (n_out <= 0 || cin <= 0 || cout <= 0 || kvol <= 0) return 0;
  if (g_wgrad_tc && msmd_spconv_bwd_weight_tc_supported(cin, cout, kvol))
    return msmd_spconv_bwd_weight_tc_workspace(n_out, cin, cout, kvol);
  return (size_t)wgrad_splits(n_out, cin, cout, kvol) * kvol * cin * cout * sizeof(float);
}

extern "C" MSMD_API int msmd_spconv_bwd_weight(const float* features, int n_in, const float* grad_out,
                                               const int* pair_fwd, int n_out, int cin, int cout,
                                               int kvol, float* grad_weight_krsc, void* workspace,
                                               size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(cin > 0 && cout > 0 && kvol > 0 && n_in >= 0 && n_out >= 0, "spconv_bwd_weight: bad sizes");
  MSMD_REQUIRE(grad_weight_krsc, "spconv_bwd_weight: null gradient");
  const size_t total = (size_t)cout * kvol * cin;
  if (n_out == 0 || n_in == 0) {
    MSMD_CUDA_OK(cudaMemsetAsync(grad_weight_krsc, 0, total * sizeof(float), stream));
    return MSMD_OK;
  }
  MSMD_REQUIRE(features && grad_out && pair_fwd, "spconv_bwd_weight: null pointer");
  if (g_wgrad_tc && msmd_spconv_bwd_weight_tc_supported(cin, cout, kvol) &&
      (((uintptr_t)features | (uintptr_t)grad_out) & 15) == 0)
    return msmd_spconv_bwd_weight_tc(features, n_in, grad_out, pair_fwd, n_out, cin, cout, kvol,
                                     grad_weight_krsc, workspace, workspace_bytes, stream_);
  const int splits = wgrad_splits(n_out, cin, cout, kvol);
  if (workspace == nullptr || workspace_bytes < (size_t)splits * total * sizeof(float)) {
    set_error("spconv_bwd_weight: workspace too small (%zu < %zu bytes)", workspace_bytes,
              (size_t)splits * total * sizeof(float));
    return MSMD_ERR_WORKSPACE;
  }
  float* partial = (float*)workspace;
  const int tiles_co = ceil_div(cout, kWgTile), tiles_ci = ceil_div(cin, kWgTile);
  dim3 grid(splits, kvol, tiles_co * tiles_ci);
  const bool vec = (cin % 4 == 0) && (cout % 4 == 0) && (((uintptr_t)features & 15) == 0) &&
                   (((uintptr_t)grad_out & 15) == 0);
  if (vec)
    spconv_wgrad_simt_kernel<true><<<grid, kWgThreads, 0, stream>>>(
        features, grad_out, pair_fwd, n_out, cin, cout, kvol, tiles_ci, splits, partial);
  else
    spconv_wgrad_simt_kernel<false><<<grid, kWgThreads, 0, stream>>>(
        features, grad_out, pair_fwd, n_out, cin, cout, kvol, tiles_ci, splits, partial);
  MSMD_LAUNCH_OK();
  spconv_wgrad_reduce_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(
      partial, splits, cout, kvol, cin, grad_weight_krsc);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_from_dense(const int* indices, const float* dense, int n, int c,
                                        int batch_size, const int* shape, float* out,
                                        msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(batch_size > 0 && c > 0 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0 && n >= 0,
               "from_dense: bad args");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(indices && dense && out, "from_dense: null pointer");
  from_dense_kernel<<<ceil_div((long long)n * c, 256), 256, 0, stream>>>(
      (const int4*)indices, dense, n, c, batch_size, shape[0], shape[1], shape[2], out);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_grid_rows(const int* indices, int n, int batch_size, const int* shape,
                                       const uint32_t* bits, const int* prefix, int* rows,
                                       msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(batch_size > 0 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0 && n >= 0,
               "grid_rows: bad args");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(indices && bits && prefix && rows, "grid_rows: null pointer");
  grid_rows_kernel<<<ceil_div(n, 256), 256, 0, stream>>>((const int4*)indices, n, batch_size, shape[0],
                                                         shape[1], shape[2], bits, prefix, rows);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
