// rulebook.cu -- occupancy bit grid + SubM / strided-conv rulebooks.
//
// spconv-2.x builds its rulebooks with a GPU hash table plus thrust sort/unique
// (ops.get_indice_pairs_implicit_gemm, call site bug_fix/conv.py:382-396).  On B200 the
// whole (batch, D, H, W) occupancy grid fits in L2 as a BITMAP (1440x1440x41 cells =
// 10.6 MB of 126 MB), so this implementation replaces hash + sort + unique by
//   bits[w]    one bit per cell, linear order ((b*D+z)*H+y)*W+x
//   prefix[w]  exclusive popcount scan over the words
// A neighbour probe is one word load + bit test (empty neighbours -- the majority --
// stop there); a hit adds one prefix load and a popcount, which yields the voxel's RANK in
// ascending linear order.  That rank IS the output row order spconv-2.x defines for strided
// convolutions (sorted unique output ids), so no sort is needed, and levels >= 2 of the
// backbone need no permutation table at all.
#include "scan.cuh"

namespace msmd {

struct Geom {
  int batch;
  int D, H, W;     // input spatial shape
  int oD, oH, oW;  // output spatial shape
  int kD, kH, kW;
  int sD, sH, sW;
  int pD, pH, pW;
  int dD, dH, dW;
};

__device__ __forceinline__ int grid_lookup(const uint32_t* __restrict__ bits,
                                           const int* __restrict__ prefix,
                                           const int* __restrict__ perm, int L) {
  const int w = L >> 5;
  const unsigned b = (unsigned)L & 31u;
  const uint32_t word = __ldg(bits + w);
  if (!((word >> b) & 1u)) return -1;
  const int r = __ldg(prefix + w) + __popc(word & ((1u << b) - 1u));
  return perm ? __ldg(perm + r) : r;
}

__global__ void __launch_bounds__(256)
grid_set_kernel(const int4* __restrict__ indices, int n, int batch, int D, int H, int W,
                uint32_t* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = indices[i];  // (b, z, y, x)
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D ||
      (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W)
    return;  // out-of-grid rows never become active cells
  const int L = ((c.x * D + c.y) * H + c.z) * W + c.w;
  atomicOr(&bits[L >> 5], 1u << (L & 31));
}

struct PopcWord {
  const uint32_t* bits;
  __device__ int operator()(int w) const { return __popc(bits[w]); }
};
struct StorePrefix {
  int* prefix;
  __device__ void operator()(int w, int ex, int) const { prefix[w] = ex; }
};

__global__ void __launch_bounds__(256)
grid_perm_kernel(const int4* __restrict__ indices, int n, int batch, int D, int H, int W,
                 const uint32_t* __restrict__ bits, const int* __restrict__ prefix,
                 int* __restrict__ perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = indices[i];
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D ||
      (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W)
    return;
  const int L = ((c.x * D + c.y) * H + c.z) * W + c.w;
  const int r = grid_lookup(bits, prefix, nullptr, L);
  atomicMax(&perm[r], i);  // duplicate coordinates: the largest row wins (deterministic)
}

// pair_fwd[k, o] for a submanifold conv: neighbour at coord(o) + (k - ksize/2) * dilation.
__global__ void __launch_bounds__(256)
subm_pairs_kernel(const int4* __restrict__ indices, int n, Geom g,
                  const uint32_t* __restrict__ bits, const int* __restrict__ prefix,
                  const int* __restrict__ perm, int* __restrict__ pair_fwd) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (o >= n) return;
  const int kz = k / (g.kH * g.kW), ky = (k / g.kW) % g.kH, kx = k % g.kW;
  const int4 c = indices[o];
  const int z = c.y + (kz - g.kD / 2) * g.dD;
  const int y = c.z + (ky - g.kH / 2) * g.dH;
  const int x = c.w + (kx - g.kW / 2) * g.dW;
  int v = -1;
  if ((unsigned)c.x < (unsigned)g.batch && (unsigned)z < (unsigned)g.D &&
      (unsigned)y < (unsigned)g.H && (unsigned)x < (unsigned)g.W)
    v = grid_lookup(bits, prefix, perm, ((c.x * g.D + z) * g.H + y) * g.W + x);
  pair_fwd[(size_t)k * n + o] = v;
}

// Strided conv, step 1: mark every output cell reachable from an active input.
//   o*s - p + k*d == i   <=>   o == (i + p - k*d) / s  when divisible and in range.
__global__ void __launch_bounds__(256)
conv_mark_kernel(const int4* __restrict__ indices, int n, Geom g, uint32_t* __restrict__ out_bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= n) return;
  const int kz = k / (g.kH * g.kW), ky = (k / g.kW) % g.kH, kx = k % g.kW;
  const int4 c = indices[i];
  if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.D ||
      (unsigned)c.z >= (unsigned)g.H || (unsigned)c.w >= (unsigned)g.W)
    return;
  const int tz = c.y + g.pD - kz * g.dD;
  const int ty = c.z + g.pH - ky * g.dH;
  const int tx = c.w + g.pW - kx * g.dW;
  if (tz < 0 || ty < 0 || tx < 0) return;
  if (tz % g.sD || ty % g.sH || tx % g.sW) return;
  const int oz = tz / g.sD, oy = ty / g.sH, ox = tx / g.sW;
  if (oz >= g.oD || oy >= g.oH || ox >= g.oW) return;
  const int L = ((c.x * g.oD + oz) * g.oH + oy) * g.oW + ox;
  atomicOr(&out_bits[L >> 5], 1u << (L & 31));
}

// Strided conv, step 2a: expand the output bitmap into (b,z,y,x) rows, ascending order.
__global__ void __launch_bounds__(256)
grid_enumerate_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ prefix,
                      int num_words, int D, int H, int W, int4* __restrict__ out_indices) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= num_words) return;
  uint32_t word = bits[w];
  if (!word) return;
  int r = prefix[w];
  while (word) {
    const int b = __ffs(word) - 1;
    word &= word - 1;
    int L = (w << 5) + b;
    int4 c;
    c.w = L % W; L /= W;
    c.z = L % H; L /= H;
    c.y = L % D; L /= D;
    c.x = L;
    out_indices[r++] = c;
  }
}

// Strided conv, step 2b: pair_fwd[k, o] = row of the active input at o*s - p + k*d.
__global__ void __launch_bounds__(256)
conv_pairs_kernel(const int4* __restrict__ out_indices, int n_out, Geom g,
                  const uint32_t* __restrict__ in_bits, const int* __restrict__ in_prefix,
                  const int* __restrict__ in_perm, int* __restrict__ pair_fwd) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (o >= n_out) return;
  const int kz = k / (g.kH * g.kW), ky = (k / g.kW) % g.kH, kx = k % g.kW;
  const int4 c = out_indices[o];
  const int z = c.y * g.sD - g.pD + kz * g.dD;
  const int y = c.z * g.sH - g.pH + ky * g.dH;
  const int x = c.w * g.sW - g.pW + kx * g.dW;
  int v = -1;
  if ((unsigned)z < (unsigned)g.D && (unsigned)y < (unsigned)g.H && (unsigned)x < (unsigned)g.W)
    v = grid_lookup(in_bits, in_prefix, in_perm, ((c.x * g.D + z) * g.H + y) * g.W + x);
  pair_fwd[(size_t)k * n_out + o] = v;
}

static int make_geom(Geom& g, int batch, const int* shape, const int* ksize, const int* stride,
                     const int* padding, const int* dilation) {
  g.batch = batch;
  g.D = shape[0]; g.H = shape[1]; g.W = shape[2];
  g.kD = ksize[0]; g.kH = ksize[1]; g.kW = ksize[2];
  g.sD = stride ? stride[0] : 1; g.sH = stride ? stride[1] : 1; g.sW = stride ? stride[2] : 1;
  g.pD = padding ? padding[0] : 0; g.pH = padding ? padding[1] : 0; g.pW = padding ? padding[2] : 0;
  g.dD = dilation ? dilation[0] : 1; g.dH = dilation ? dilation[1] : 1; g.dW = dilation ? dilation[2] : 1;
  MSMD_REQUIRE(batch > 0 && g.D > 0 && g.H > 0 && g.W > 0, "rulebook: empty grid");
  MSMD_REQUIRE(g.kD > 0 && g.kH > 0 && g.kW > 0 && g.sD > 0 && g.sH > 0 && g.sW > 0 && g.dD > 0 &&
                   g.dH > 0 && g.dW > 0 && g.pD >= 0 && g.pH >= 0 && g.pW >= 0,
               "rulebook: bad conv geometry");
  g.oD = (g.D + 2 * g.pD - g.dD * (g.kD - 1) - 1) / g.sD + 1;
  g.oH = (g.H + 2 * g.pH - g.dH * (g.kH - 1) - 1) / g.sH + 1;
  g.oW = (g.W + 2 * g.pW - g.dW * (g.kW - 1) - 1) / g.sW + 1;
  MSMD_REQUIRE((long long)batch * g.D * g.H * g.W < 0x7fffffffLL,
               "rulebook: batch*D*H*W exceeds the int32 cell index");
  return MSMD_OK;
}

static size_t num_words(long long cells) { return (size_t)((cells + 31) / 32); }

struct ScanWs {
  int* block_sums;
  int* total;
  unsigned* counter;
  bool carve(Workspace& ws) {
    block_sums = ws.take<int>(kScanMaxBlocks);
    total = ws.take<int>(1);
    counter = ws.take<unsigned>(1);
    return ws.ok();
  }
};

static int scan_bitmap(uint32_t* bits, int* prefix, size_t words, int* total_out, void* workspace,
                       size_t workspace_bytes, cudaStream_t stream) {
  Workspace ws(workspace, workspace_bytes);
  ScanWs s;
  if (!s.carve(ws)) {
    set_error("rulebook: scan workspace too small (%zu < %zu)", workspace_bytes,
              msmd_scan_workspace());
    return MSMD_ERR_WORKSPACE;
  }
  MSMD_REQUIRE(words < 0x7fffffffULL, "rulebook: bitmap too large");
  MSMD_CUDA_OK(cudaMemsetAsync(s.counter, 0, sizeof(unsigned), stream));
  ScanTemp<int> tmp{s.block_sums, s.counter, total_out ? total_out : s.total};
  MSMD_CUDA_OK((device_exclusive_scan<int>(PopcWord{bits}, StorePrefix{prefix}, (int)words, tmp,
                                           stream)));
  return MSMD_OK;
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API size_t msmd_grid_num_words(int batch_size, const int* s) {
  return num_words((long long)batch_size * s[0] * s[1] * s[2]);
}

extern "C" MSMD_API size_t msmd_scan_workspace(void) {
  Workspace ws((void*)256, ~(size_t)0 >> 1);
  ScanWs s;
  s.carve(ws);
  return ws.used + 256;
}

extern "C" MSMD_API int msmd_conv_out_shape(const int* shape, const int* ksize, const int* stride,
                                   const int* padding, const int* dilation, int* out_shape) {
  Geom g;
  int r = make_geom(g, 1, shape, ksize, stride, padding, dilation);
  if (r) return r;
  out_shape[0] = g.oD; out_shape[1] = g.oH; out_shape[2] = g.oW;
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_grid_build(const int* indices, int n, int batch_size, const int* shape,
                               uint32_t* bits, int* prefix, int* perm, int* num_active,
                               void* workspace, size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int one[3] = {1, 1, 1};
  Geom g;
  int r = make_geom(g, batch_size, shape, one, nullptr, nullptr, nullptr);
  if (r) return r;
  MSMD_REQUIRE(n >= 0 && bits && prefix, "grid_build: null argument");
  const size_t words = msmd_grid_num_words(batch_size, shape);
  MSMD_CUDA_OK(cudaMemsetAsync(bits, 0, words * sizeof(uint32_t), stream));
  if (n > 0) {
    grid_set_kernel<<<ceil_div(n, 256), 256, 0, stream>>>((const int4*)indices, n, batch_size, g.D,
                                                          g.H, g.W, bits);
    MSMD_LAUNCH_OK();
  }
  r = scan_bitmap(bits, prefix, words, num_active, workspace, workspace_bytes, stream);
  if (r) return r;
  if (perm && n > 0) {
    MSMD_CUDA_OK(cudaMemsetAsync(perm, 0xFF, (size_t)n * sizeof(int), stream));
    grid_perm_kernel<<<ceil_div(n, 256), 256, 0, stream>>>((const int4*)indices, n, batch_size, g.D,
                                                           g.H, g.W, bits, prefix, perm);
    MSMD_LAUNCH_OK();
  }
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_rulebook_subm(const int* indices, int n, int batch_size, const int* shape,
                                  const int* ksize, const int* dilation, const uint32_t* bits,
                                  const int* prefix, const int* perm, int* pair_fwd,
                                  msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Geom g;
  int r = make_geom(g, batch_size, shape, ksize, nullptr, nullptr, dilation);
  if (r) return r;
  if (n == 0) return MSMD_OK;
  const int kvol = g.kD * g.kH * g.kW;
  MSMD_REQUIRE(kvol <= 65535, "rulebook_subm: kernel volume too large");
  dim3 grid(ceil_div(n, 256), kvol);
  subm_pairs_kernel<<<grid, 256, 0, stream>>>((const int4*)indices, n, g, bits, prefix, perm,
                                              pair_fwd);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_rulebook_conv_outputs(const int* indices, int n, int batch_size,
                                          const int* shape, const int* ksize, const int* stride,
                                          const int* padding, const int* dilation,
                                          uint32_t* out_bits, int* out_prefix, int* num_out,
                                          void* workspace, size_t workspace_bytes,
                                          msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Geom g;
  int r = make_geom(g, batch_size, shape, ksize, stride, padding, dilation);
  if (r) return r;
  MSMD_REQUIRE(g.oD > 0 && g.oH > 0 && g.oW > 0, "rulebook_conv: empty output shape");
  const int oshape[3] = {g.oD, g.oH, g.oW};
  const size_t words = msmd_grid_num_words(batch_size, oshape);
  MSMD_CUDA_OK(cudaMemsetAsync(out_bits, 0, words * sizeof(uint32_t), stream));
  const int kvol = g.kD * g.kH * g.kW;
  MSMD_REQUIRE(kvol <= 65535, "rulebook_conv: kernel volume too large");
  if (n > 0) {
    dim3 grid(ceil_div(n, 256), kvol);
    conv_mark_kernel<<<grid, 256, 0, stream>>>((const int4*)indices, n, g, out_bits);
    MSMD_LAUNCH_OK();
  }
  return scan_bitmap(out_bits, out_prefix, words, num_out, workspace, workspace_bytes, stream);
}

extern "C" MSMD_API int msmd_rulebook_conv_pairs(const uint32_t* out_bits, const int* out_prefix, int n_out,
                                        int batch_size, const int* shape, const int* ksize,
                                        const int* stride, const int* padding, const int* dilation,
                                        const uint32_t* in_bits, const int* in_prefix,
                                        const int* in_perm, int* out_indices, int* pair_fwd,
                                        msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  Geom g;
  int r = make_geom(g, batch_size, shape, ksize, stride, padding, dilation);
  if (r) return r;
  if (n_out == 0) return MSMD_OK;
  const int oshape[3] = {g.oD, g.oH, g.oW};
  const size_t words = msmd_grid_num_words(batch_size, oshape);
  grid_enumerate_kernel<<<ceil_div((long long)words, 256), 256, 0, stream>>>(
      out_bits, out_prefix, (int)words, g.oD, g.oH, g.oW, (int4*)out_indices);
  MSMD_LAUNCH_OK();
  const int kvol = g.kD * g.kH * g.kW;
  dim3 grid(ceil_div(n_out, 256), kvol);
  conv_pairs_kernel<<<grid, 256, 0, stream>>>((const int4*)out_indices, n_out, g, in_bits,
                                              in_prefix, in_perm, pair_fwd);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
