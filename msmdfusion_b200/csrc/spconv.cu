// spconv.cu -- sparse convolution forward (gather -> contraction -> fused epilogue) and
// SparseConvTensor.dense().
//
//   out[o, co] = sum_k sum_ci  x[pair_fwd[k, o], ci] * W[co, k, ci]      (fp32 accumulate)
//
// Replaces spconv-2.x Fsp.implicit_gemm (call site bug_fix/conv.py:442-447).  This file is
// the exact-fp32 path (FFMA): one CTA owns a tile of TM output voxels x TN output channels,
// walks the kernel offsets k (skipping offsets no voxel of the tile uses), gathers the
// TM x 16 input slab for that offset into shared memory (transposed, conflict-free) and
// multiplies it with the 16 x TN weight slab.  BatchNorm(eval) scale/shift, the residual
// add of SparseBasicBlock and ReLU are fused into the store, so a conv+BN+ReLU layer
// moves  4*(N_in*Cin + N_out*Cout + K*Cin*Cout) + 4*K*N_out  bytes once.
#include "common.cuh"

namespace msmd {

constexpr int kConvThreads = 256;
constexpr int kConvKC = 16;

template <int RM, int RN, bool VEC4>
__global__ void __launch_bounds__(kConvThreads)
spconv_fwd_simt_kernel(const float* __restrict__ feat, const float* __restrict__ wp,
                       const int* __restrict__ pair, int n_out, int cin, int cout, int kvol,
                       const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ residual, int relu, float* __restrict__ out) {
  constexpr int TM = 16 * RM, TN = 16 * RN, KC = kConvKC;
  __shared__ float A_s[KC][TM + 2];
  __shared__ __align__(16) float B_s[KC][TN];
  __shared__ int idx_s[TM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int row0 = blockIdx.x * TM;
  const int col0 = blockIdx.y * TN;

  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < kvol; ++k) {
    int has = 0;
    if (tid < TM) {
      const int o = row0 + tid;
      const int p = (o < n_out) ? __ldg(pair + (size_t)k * n_out + o) : -1;
      idx_s[tid] = p;
      has = p >= 0;
    }
    if (!__syncthreads_or(has)) continue;  // no voxel of this tile uses offset k

    const float* wk = wp + (size_t)k * cin * cout;
    for (int ci0 = 0; ci0 < cin; ci0 += KC) {
      // ---- gather the TM x KC input slab (zero rows where pair == -1) ----
      if (VEC4) {
        for (int s = tid; s < TM * 4; s += kConvThreads) {
          const int r = s >> 2, q = s & 3;
          const int p = idx_s[r];
          const int c = ci0 + q * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p >= 0 && c < cin) v = __ldg((const float4*)(feat + (size_t)p * cin + c));
          A_s[q * 4 + 0][r] = v.x;
          A_s[q * 4 + 1][r] = v.y;
          A_s[q * 4 + 2][r] = v.z;
          A_s[q * 4 + 3][r] = v.w;
        }
      } else {
        for (int s = tid; s < TM * KC; s += kConvThreads) {
          const int r = s / KC, kk = s % KC;
          const int p = idx_s[r];
          const int c = ci0 + kk;
          A_s[kk][r] = (p >= 0 && c < cin) ? __ldg(feat + (size_t)p * cin + c) : 0.f;
        }
      }
      // ---- weight slab KC x TN ----
      for (int s = tid; s < KC * TN; s += kConvThreads) {
        const int kk = s / TN, j = s % TN;
        const int c = ci0 + kk, co = col0 + j;
        B_s[kk][j] = (c < cin && co < cout) ? __ldg(wk + (size_t)c * cout + co) : 0.f;
      }
      __syncthreads();
      const int kc = min(KC, cin - ci0);
#pragma unroll 4
      for (int kk = 0; kk < kc; ++kk) {
        float a[RM], b[RN];
#pragma unroll
        for (int i = 0; i < RM; ++i) a[i] = A_s[kk][ty + 16 * i];
        if constexpr (RN == 4) {
          const float4 t = *(const float4*)&B_s[kk][tx * 4];
          b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
        } else if constexpr (RN == 2) {
          const float2 t = *(const float2*)&B_s[kk][tx * 2];
          b[0] = t.x; b[1] = t.y;
        } else {
          b[0] = B_s[kk][tx];
        }
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
          for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- fused epilogue: BN(eval) scale/shift, residual, ReLU ----
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int o = row0 + ty + 16 * i;
    if (o >= n_out) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int co = col0 + tx * RN + j;
      if (co >= cout) continue;
      float v = acc[i][j];
      if (scale) v = fmaf(v, __ldg(scale + co), __ldg(shift + co));
      if (residual) v += __ldg(residual + (size_t)o * cout + co);
      if (relu) v = fmaxf(v, 0.f);
      out[(size_t)o * cout + co] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, int cout, int kvol, int cin,
                   float* __restrict__ packed) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)cout * kvol * cin;
  if (t >= total) return;
  // t enumerates the packed layout [k][ci][co]
  const int co = (int)(t % cout);
  const int ci = (int)((t / cout) % cin);
  const int k = (int)(t / ((size_t)cout * cin));
  packed[t] = w[((size_t)co * kvol + k) * cin + ci];
}

__global__ void __launch_bounds__(256)
to_dense_kernel(const int4* __restrict__ indices, const float* __restrict__ feat, int n, int C,
                int batch, int D, int H, int W, float* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * C) return;
  const int i = (int)(t % n), c = (int)(t / n);
  const int4 p = indices[i];
  if ((unsigned)p.x >= (unsigned)batch || (unsigned)p.y >= (unsigned)D ||
      (unsigned)p.z >= (unsigned)H || (unsigned)p.w >= (unsigned)W)
    return;
  out[((((size_t)p.x * C + c) * D + p.y) * H + p.z) * W + p.w] = feat[(size_t)i * C + c];
}

template <int RM, int RN>
static cudaError_t launch_simt(const float* feat, const float* wp, const int* pair, int n_out,
                               int cin, int cout, int kvol, const float* scale,
                               const float* shift, const float* residual, int relu, float* out,
                               cudaStream_t stream) {
  dim3 grid(ceil_div(n_out, 16 * RM), ceil_div(cout, 16 * RN));
  const bool vec = (cin % 4 == 0) && (((uintptr_t)feat & 15) == 0);
  if (vec)
    spconv_fwd_simt_kernel<RM, RN, true><<<grid, kConvThreads, 0, stream>>>(
        feat, wp, pair, n_out, cin, cout, kvol, scale, shift, residual, relu, out);
  else
    spconv_fwd_simt_kernel<RM, RN, false><<<grid, kConvThreads, 0, stream>>>(
        feat, wp, pair, n_out, cin, cout, kvol, scale, shift, residual, relu, out);
  return cudaGetLastError();
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API int msmd_spconv_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                       float* packed, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(weight_krsc && packed && cout > 0 && kvol > 0 && cin > 0, "pack_weight: bad args");
  const size_t total = (size_t)cout * kvol * cin;
  pack_weight_kernel<<<ceil_div((long long)total, 256), 256, 0, stream>>>(weight_krsc, cout, kvol,
                                                                          cin, packed);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_fwd(const float* features, int n_in, const float* packed_weight,
                               const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                               const float* scale, const float* shift, const float* residual,
                               int relu, float* out, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(cin > 0 && cout > 0 && kvol > 0 && n_in >= 0 && n_out >= 0, "spconv_fwd: bad sizes");
  MSMD_REQUIRE((scale == nullptr) == (shift == nullptr), "spconv_fwd: scale/shift must come together");
  if (n_out == 0) return MSMD_OK;
  MSMD_REQUIRE(features && packed_weight && pair_fwd && out, "spconv_fwd: null pointer");
  // tile choice: wide channel tiles for wide layers; tall row tiles only when there are
  // enough output voxels to keep >= 4 CTAs per SM busy.
  const int ytiles64 = ceil_div(cout, 64);
  const bool tall = (long long)ceil_div(n_out, 128) * ytiles64 >= 4LL * kNumSMs;
  cudaError_t e;
#define MSMD_SIMT(RM, RN)                                                                        \
  e = launch_simt<RM, RN>(features, packed_weight, pair_fwd, n_out, cin, cout, kvol, scale, shift, \
                          residual, relu, out, stream)
  if (cout <= 16) {
    if (tall) MSMD_SIMT(8, 1); else MSMD_SIMT(4, 1);
  } else if (cout <= 32) {
    if (tall) MSMD_SIMT(8, 2); else MSMD_SIMT(4, 2);
  } else {
    if (tall) MSMD_SIMT(8, 4); else MSMD_SIMT(4, 4);
  }
#undef MSMD_SIMT
  MSMD_CUDA_OK(e);
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_to_dense(const int* indices, const float* features, int n, int c,
                             int batch_size, const int* shape, float* out, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(out && batch_size > 0 && c > 0 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0,
               "to_dense: bad args");
  const size_t total = (size_t)batch_size * c * shape[0] * shape[1] * shape[2];
  MSMD_CUDA_OK(cudaMemsetAsync(out, 0, total * sizeof(float), stream));
  if (n == 0) return MSMD_OK;
  const size_t work = (size_t)n * c;
  to_dense_kernel<<<ceil_div((long long)work, 256), 256, 0, stream>>>(
      (const int4*)indices, features, n, c, batch_size, shape[0], shape[1], shape[2], out);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
