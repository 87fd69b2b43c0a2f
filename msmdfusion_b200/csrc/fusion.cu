// fusion.cu -- the index/gather ops either side of the Gated Modality-Aware convolution:
//   * voxel_modality_split   (MSMDFusion.py:27-45, :251-325)
//   * Fsp.sparse_add         (call site sparse_multimodal_encoder_painting.py:455)
//   * get_foreground2D lift  (MSMDFusion.py:169-238): pixel-feature gather x score gate
#include "sort.cuh"

namespace msmd {

// ------------------------------------------------------------------------------------
// voxel_modality_split
// The reference builds a FLOAT32 key z*1e6 + y*1e3 + x (int32 tensor * python float ->
// float32; MSMDFusion.py:271-272), sorts both voxel sets by it, and pairs equal keys
// one-to-one with a CPU two-pointer merge (type_assign).  Keys above 2^24 collide for
// neighbouring x -- that rounding is part of the reference's behaviour and is reproduced
// exactly (separate fp32 mul, mul, add, add; no FMA).  The keys are exact integers < 2^26, so
// they sort as uint32; the sort is stable (ties keep row order).  For a run of `a` equal keys
// in one set and `b` in the other the merge flags the first min(a,b) of each:
//        flagged(i)  <=>  (i - lower_bound_own(key)) < count_other(key).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(int z, int y, int x) {
  const float a = __fmul_rn((float)z, 1e6f);
  const float b = __fmul_rn((float)y, 1e3f);
  const float k = __fadd_rn(__fadd_rn(a, b), (float)x);
  return (uint32_t)k;
}

__global__ void __launch_bounds__(256)
split_keys_kernel(const int* __restrict__ coord, int stride, int n, uint32_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* c = coord + (size_t)i * stride;  // (b, z, y, x)
  keys[i] = float_key(c[1], c[2], c[3]);
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound_u32(const uint32_t* a, int n, uint32_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
split_flag_kernel(const uint32_t* __restrict__ own, const int* __restrict__ own_rows, int n_own,
                  const uint32_t* __restrict__ other, int n_other, int* __restrict__ flag_sorted,
                  int* __restrict__ mix_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_own) return;
  const uint32_t key = own[i];
  const int j = i - lower_bound_u32(own, n_own, key);
  const int cnt = upper_bound_u32(other, n_other, key) - lower_bound_u32(other, n_other, key);
  const int f = j < cnt;
  flag_sorted[i] = f;
  mix_rows[own_rows[i]] = f;
}

// ---- hash path (default) -----------------------------------------------------------------------------
// The sort above costs ~35 launches (two 4-pass radix sorts, binary-search flags, scans) = 0.26 ms per call, all of
// it on the serial path in front of the FPS chains.  What the merge needs is (a) for every row the number of rows of
// the OTHER set with the same key, (b) its rank among the rows of its OWN set with that key (runs are 1..5 rows:
// keys collide only through fp32 rounding above 2^24 and through x >= 1000), (c) the flagged rows in sorted-key
// order.  (a) and (b) come from an open-addressing table keyed by the float key: a slot counts the rows of either
// set and remembers up to kRunCap row ids per set, in whatever order the atomics land -- the rank is the number of
// remembered ids BELOW the row's own, which does not depend on that order.  (c) only concerns the flagged rows (the
// voxels present in both modalities, a few hundred): they are compacted and sorted by (key, row) in shared memory by
// one CTA per set.  Five launches.  A run longer than kRunCap or more than kSplitSortCap pairs raise `overflow` and
// the caller falls back to the sort path.
constexpr int kRunCap = 8;
constexpr int kSplitSortCap = 4096;   // 32 KB of static shared memory
constexpr uint32_t kEmptyKey = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t split_hash(uint32_t k) {
  k ^= k >> 16; k *= 0x7feb352dU; k ^= k >> 15; k *= 0x846ca68bU; k ^= k >> 16;
  return k;
}

__global__ void __launch_bounds__(256)
split_insert_kernel(const int* __restrict__ coord3, int n3, const int* __restrict__ coord2, int n2, uint32_t mask,
                    uint32_t* __restrict__ keys, int* __restrict__ cnt /* [2][cap] */, int* __restrict__ ids /* [2][cap][kRunCap] */,
                    int* __restrict__ slot_of /* [n3 + n2] */, uint32_t* __restrict__ key_of, int* __restrict__ overflow) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n3 + n2) return;
  const int set = t >= n3;
  const int row = set ? t - n3 : t;
  const int* c = (set ? coord2 : coord3) + (size_t)row * 4;
  const uint32_t key = float_key(c[1], c[2], c[3]);
  uint32_t h = split_hash(key) & mask;
  for (;;) {
    const uint32_t old = atomicCAS(&keys[h], kEmptyKey, key);
    if (old == kEmptyKey || old == key) break;
    h = (h + 1) & mask;
  }
  const size_t cap = (size_t)mask + 1;
  const int pos = atomicAdd(&cnt[set * cap + h], 1);
  if (pos < kRunCap) ids[((size_t)set * cap + h) * kRunCap + pos] = row;
  else atomicExch(overflow, 1);
  slot_of[t] = (int)h;
  key_of[t] = key;
}

__global__ void __launch_bounds__(256)
split_flag_hash_kernel(int n3, int n2, uint32_t mask, const int* __restrict__ cnt, const int* __restrict__ ids,
                       const int* __restrict__ slot_of, const uint32_t* __restrict__ key_of, int* __restrict__ mix3,
                       int* __restrict__ mix2, unsigned long long* __restrict__ flagged /* [2][min(n3,n2)] */, int cap_pairs,
                       int* __restrict__ num_flag /* [2] */, int* __restrict__ overflow) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n3 + n2) return;
  const int set = t >= n3;
  const int row = set ? t - n3 : t;
  const size_t cap = (size_t)mask + 1;
  const int h = slot_of[t];
  int own = cnt[set * cap + h];
  if (own > kRunCap) own = kRunCap;
  const int other = cnt[(1 - set) * cap + h];
  int rank = 0;
  const int* run = ids + ((size_t)set * cap + h) * kRunCap;
  for (int j = 0; j < own; ++j) rank += run[j] < row;
  const int f = rank < other;
  (set ? mix2 : mix3)[row] = f;
  if (f) {
    const int p = atomicAdd(&num_flag[set], 1);
    if (p < cap_pairs) flagged[(size_t)set * cap_pairs + p] = ((unsigned long long)key_of[t] << 32) | (unsigned)row;
    else atomicExch(overflow, 1);
  }
}

// one CTA per set: bitonic sort of the flagged (key, row) pairs in shared memory -> syn list (+ offset)
__global__ void __launch_bounds__(1024)
split_sort_flagged_kernel(const unsigned long long* __restrict__ flagged, int cap_pairs, const int* __restrict__ num_flag,
                          long long offset3, long long offset2, long long* __restrict__ syn3,
                          long long* __restrict__ syn2, int* __restrict__ num_mix, int* __restrict__ overflow) {
  __shared__ unsigned long long v[kSplitSortCap];
  const int set = blockIdx.x;
  const int p = num_flag[set];
  if (set == 0 && threadIdx.x == 0) *num_mix = p;
  if (p > kSplitSortCap || p > cap_pairs || num_flag[0] != num_flag[1]) {
    if (threadIdx.x == 0 && (p > kSplitSortCap || p > cap_pairs)) atomicExch(overflow, 1);
    if (num_flag[0] != num_flag[1] && threadIdx.x == 0) atomicExch(overflow, 1);   // cannot happen: caught, not trusted
    return;
  }
  int n = 1;
  while (n < p) n <<= 1;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = i < p ? flagged[(size_t)set * cap_pairs + i] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = v[i], b = v[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { v[i] = b; v[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  long long* syn = set ? syn2 : syn3;
  const long long off = set ? offset2 : offset3;
  for (int i = threadIdx.x; i < p; i += blockDim.x) syn[i] = (long long)(unsigned)(v[i] & 0xffffffffull) + off;
}

struct SplitHashWs {
  uint32_t* keys;
  int *cnt, *ids, *slot_of, *num_flag;
  uint32_t* key_of;
  unsigned long long* flagged;
  uint32_t cap;
  int cap_pairs;
  bool carve(Workspace& ws, int n3, int n2) {
    const long long n = (long long)n3 + n2;
    cap = 1024;
    while ((long long)cap < 2 * n) cap <<= 1;
    cap_pairs = n3 < n2 ? n3 : n2;
    if (cap_pairs < 1) cap_pairs = 1;
    keys = ws.take<uint32_t>(cap);
    cnt = ws.take<int>((size_t)2 * cap);
    ids = ws.take<int>((size_t)2 * cap * kRunCap);
    slot_of = ws.take<int>(n > 0 ? n : 1);
    key_of = ws.take<uint32_t>(n > 0 ? n : 1);
    flagged = ws.take<unsigned long long>((size_t)2 * cap_pairs);
    num_flag = ws.take<int>(2);
    return ws.ok();
  }
};

struct SplitCompact {
  const int* rows;
  long long offset;
  long long* out;
  __device__ void operator()(int i, int ex, int v) const {
    if (v) out[ex] = (long long)rows[i] + offset;
  }
};

struct LoadUnflagged {
  const int* f;
  __device__ int operator()(int i) const { return f[i] == 0; }
};
struct StoreRow {
  long long* out;
  __device__ void operator()(int i, int ex, int v) const {
    if (v) out[ex] = (long long)i;
  }
};

struct SplitWs {
  uint32_t *k3, *k2;
  int *r3, *r2, *f3, *f2;
  int* total2;
  SortWs sort;
  bool carve(Workspace& ws, int n3, int n2) {
    k3 = ws.take<uint32_t>(n3 > 0 ? n3 : 1);
    k2 = ws.take<uint32_t>(n2 > 0 ? n2 : 1);
    r3 = ws.take<int>(n3 > 0 ? n3 : 1);
    r2 = ws.take<int>(n2 > 0 ? n2 : 1);
    f3 = ws.take<int>(n3 > 0 ? n3 : 1);
    f2 = ws.take<int>(n2 > 0 ? n2 : 1);
    total2 = ws.take<int>(1);
    return sort.carve(ws, n3 > n2 ? n3 : n2);
  }
};

// ------------------------------------------------------------------------------------
// sparse_add: union of two sparse tensors, features of coincident voxels summed, output
// rows ascending by linear index (spconv v2.1.21: torch.sparse coalesce).  The union is the
// OR of both occupancy bitmaps; the popcount scan yields the sorted output rows directly,
// and the same grid serves the strided convolution that follows.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_or_kernel(const int4* __restrict__ indices, int n, int batch, int D, int H, int W,
               uint32_t* __restrict__ bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 c = indices[i];
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D ||
      (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W)
    return;
  const int L = ((c.x * D + c.y) * H + c.z) * W + c.w;
  atomicOr(&bits[L >> 5], 1u << (L & 31));
}

struct PopcWord2 {
  const uint32_t* bits;
  __device__ int operator()(int w) const { return __popc(bits[w]); }
};
struct StorePrefix2 {
  int* prefix;
  __device__ void operator()(int w, int ex, int) const { prefix[w] = ex; }
};

__global__ void __launch_bounds__(256)
union_enumerate_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ prefix,
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

__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const int4* __restrict__ indices, const float* __restrict__ feat, int n,
                        int C, int batch, int D, int H, int W, const uint32_t* __restrict__ bits,
                        const int* __restrict__ prefix, float* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n * C) return;
  const int i = (int)(t / C), ch = (int)(t % C);
  const int4 c = indices[i];
  if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)D ||
      (unsigned)c.z >= (unsigned)H || (unsigned)c.w >= (unsigned)W)
    return;
  const int L = ((c.x * D + c.y) * H + c.z) * W + c.w;
  const int w = L >> 5;
  const unsigned b = (unsigned)L & 31u;
  const uint32_t word = __ldg(bits + w);
  const int r = __ldg(prefix + w) + __popc(word & ((1u << b) - 1u));
  atomicAdd(out + (size_t)r * C + ch, feat[t]);
}

// ------------------------------------------------------------------------------------
// Lift: per virtual point, gather the C-channel image feature at its pixel, gate it with
// score = ReLU(Linear_{C+17 -> 1}([feat, depth, lidar2img(16)])) and emit
// [point(P) | feat * score]  (MSMDFusion.py:202-230).  One warp per point; the feature map is
// addressed through explicit strides so a channels-last map gives 196-byte coalesced reads.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lift_gather_kernel(const float* __restrict__ img, long long s_cam, long long s_c, long long s_y,
                   long long s_x, int C, int H, int W, const float* __restrict__ pix,
                   const int* __restrict__ cam, const float* __restrict__ pts, int P, int M,
                   const float* __restrict__ lidar2img, float downscale,
                   const float* __restrict__ score_w, float score_b, float* __restrict__ out) {
  const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (p >= M) return;
  const float u = pix[3 * p + 0], v = pix[3 * p + 1], depth = pix[3 * p + 2];
  // fg_pxl * downscale_factor in fp32, then .long() (truncation toward zero), :207-209
  int cw = (int)truncf(__fmul_rn(u, downscale));
  int chh = (int)truncf(__fmul_rn(v, downscale));
  cw = min(max(cw, 0), W - 1);  // the reference would raise on out-of-map pixels
  chh = min(max(chh, 0), H - 1);
  const int cm = cam[p];
  const float* base = img + (long long)cm * s_cam + (long long)chh * s_y + (long long)cw * s_x;
  float f[4];
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    f[j] = 0.f;
    if (c < C) {
      f[j] = __ldg(base + (long long)c * s_c);
      acc = fmaf(f[j], __ldg(score_w + c), acc);
    }
  }
  if (lane == 0) acc = fmaf(depth, __ldg(score_w + C), acc);
  if (lane < 16) acc = fmaf(__ldg(lidar2img + cm * 16 + lane), __ldg(score_w + C + 1 + lane), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const float score = fmaxf(acc + score_b, 0.f);
  float* o = out + (size_t)p * (P + C);
  for (int c = lane; c < P; c += 32) o[c] = pts[(size_t)p * P + c];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = lane + 32 * j;
    if (c < C) o[P + c] = f[j] * score;
  }
}


// ------------------------------------------------------------------------------------
// Mask-sorted tiles (the role of spconv-2.x's mask_argsort_fwd_splits, bug_fix/conv.py:382-415).
// The tensor-core convolution skips a K chunk only when NO row of a 128-row tile uses its kernel
// offsets; with rows in index-set order a tile touches 23-27 of the 27 offsets although a row uses
// 12-50 % of them.  Rows are therefore grouped by a 15-bit STRUCTURAL digest of their neighbour mask
// (k = (kz*3+ky)*3+kx): the nine bits of the voxel's own z plane, one "any neighbour in this ky row"
// bit per row of the plane above, then of the plane below -- LiDAR surfaces are thin sheets, most
// voxels have nothing above / below.  Stable 2-pass radix sort (sort.cuh), then the pair table is
// permuted once; the convolution reads the permuted table and writes row row_perm[slot].
// Measured on the CPU (profiles/r01g_mask_sort_estimate.txt): 4.4x / 2.0x / 1.7x / 1.2x fewer
// tile x offset products at the four resolution levels.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_digest_kernel(const int* __restrict__ pair, int n, uint32_t* __restrict__ keys) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n) return;
  uint32_t key = 0;
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const bool used = __ldg(pair + (size_t)k * n + o) >= 0;  // coalesced along o
    if (!used) continue;
    const int kz = k / 9, ky = (k / 3) % 3;
    if (kz == 1) key |= 1u << (k - 9);       // own plane: bits 0..8
    else if (kz == 2) key |= 1u << (9 + ky);  // plane above: bits 9..11
    else key |= 1u << (12 + ky);              // plane below: bits 12..14
  }
  keys[o] = key;
}

__global__ void __launch_bounds__(256)
pair_permute_kernel(const int* __restrict__ pair, const int* __restrict__ row_perm, int kvol, int n,
                    int* __restrict__ pair_sorted) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)kvol * n) return;
  const int k = (int)(t / n), slot = (int)(t % n);
  pair_sorted[t] = __ldg(pair + (size_t)k * n + __ldg(row_perm + slot));
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API size_t msmd_modality_split_workspace(int n3, int n2) {
  Workspace ws((void*)256, ~(size_t)0 >> 1);
  SplitWs s;
  s.carve(ws, n3, n2);
  Workspace wh((void*)256, ~(size_t)0 >> 1);
  SplitHashWs h;
  h.carve(wh, n3, n2);
  return (ws.used > wh.used ? ws.used : wh.used) + 256;
}

// Hash path (see above).  num_mix[0] = number of pairs, num_mix[1] = 1 when the table / pair buffers overflowed: the
// outputs are then undefined and the caller re-runs msmd_modality_split_sort (same arguments).
extern "C" MSMD_API int msmd_modality_split(const int* coord3, int n3, const int* coord2, int n2,
                                            long long offset3, long long offset2, int* mix3,
                                            int* mix2, long long* syn3, long long* syn2,
                                            int* num_mix, void* workspace, size_t workspace_bytes,
                                            msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(n3 >= 0 && n2 >= 0 && num_mix, "modality_split: bad arguments");
  MSMD_CUDA_OK(cudaMemsetAsync(num_mix, 0, 2 * sizeof(int), stream));
  if (n3 == 0 || n2 == 0) {
    if (n3) MSMD_CUDA_OK(cudaMemsetAsync(mix3, 0, (size_t)n3 * sizeof(int), stream));
    if (n2) MSMD_CUDA_OK(cudaMemsetAsync(mix2, 0, (size_t)n2 * sizeof(int), stream));
    return MSMD_OK;
  }
  Workspace ws(workspace, workspace_bytes);
  SplitHashWs h;
  if (!h.carve(ws, n3, n2)) {
    set_error("modality_split: workspace too small (%zu < %zu)", workspace_bytes,
              msmd_modality_split_workspace(n3, n2));
    return MSMD_ERR_WORKSPACE;
  }
  MSMD_CUDA_OK(cudaMemsetAsync(h.keys, 0xFF, (size_t)h.cap * sizeof(uint32_t), stream));
  MSMD_CUDA_OK(cudaMemsetAsync(h.cnt, 0, (size_t)2 * h.cap * sizeof(int), stream));
  MSMD_CUDA_OK(cudaMemsetAsync(h.num_flag, 0, 2 * sizeof(int), stream));
  const int n = n3 + n2;
  split_insert_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(coord3, n3, coord2, n2, h.cap - 1, h.keys, h.cnt, h.ids,
                                                           h.slot_of, h.key_of, num_mix + 1);
  MSMD_LAUNCH_OK();
  split_flag_hash_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(n3, n2, h.cap - 1, h.cnt, h.ids, h.slot_of, h.key_of,
                                                              mix3, mix2, h.flagged, h.cap_pairs, h.num_flag,
                                                              num_mix + 1);
  MSMD_LAUNCH_OK();
  split_sort_flagged_kernel<<<2, 1024, 0, stream>>>(h.flagged, h.cap_pairs, h.num_flag, offset3, offset2, syn3, syn2,
                                                    num_mix, num_mix + 1);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_modality_split_sort(const int* coord3, int n3, const int* coord2, int n2,
                                            long long offset3, long long offset2, int* mix3,
                                            int* mix2, long long* syn3, long long* syn2,
                                            int* num_mix, void* workspace, size_t workspace_bytes,
                                            msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(n3 >= 0 && n2 >= 0 && num_mix, "modality_split: bad arguments");
  Workspace ws(workspace, workspace_bytes);
  SplitWs s;
  if (!s.carve(ws, n3, n2)) {
    set_error("modality_split: workspace too small (%zu < %zu)", workspace_bytes,
              msmd_modality_split_workspace(n3, n2));
    return MSMD_ERR_WORKSPACE;
  }
  if (n3 == 0 || n2 == 0) {
    if (n3) MSMD_CUDA_OK(cudaMemsetAsync(mix3, 0, (size_t)n3 * sizeof(int), stream));
    if (n2) MSMD_CUDA_OK(cudaMemsetAsync(mix2, 0, (size_t)n2 * sizeof(int), stream));
    MSMD_CUDA_OK(cudaMemsetAsync(num_mix, 0, sizeof(int), stream));
    return MSMD_OK;
  }
  split_keys_kernel<<<ceil_div(n3, 256), 256, 0, stream>>>(coord3, 4, n3, s.k3);
  MSMD_LAUNCH_OK();
  split_keys_kernel<<<ceil_div(n2, 256), 256, 0, stream>>>(coord2, 4, n2, s.k2);
  MSMD_LAUNCH_OK();
  MSMD_CUDA_OK(radix_sort_pairs(s.k3, s.r3, n3, 26, true, s.sort, stream));
  MSMD_CUDA_OK(radix_sort_pairs(s.k2, s.r2, n2, 26, true, s.sort, stream));
  split_flag_kernel<<<ceil_div(n3, 256), 256, 0, stream>>>(s.k3, s.r3, n3, s.k2, n2, s.f3, mix3);
  MSMD_LAUNCH_OK();
  split_flag_kernel<<<ceil_div(n2, 256), 256, 0, stream>>>(s.k2, s.r2, n2, s.k3, n3, s.f2, mix2);
  MSMD_LAUNCH_OK();
  MSMD_CUDA_OK(cudaMemsetAsync(s.sort.counter, 0, sizeof(unsigned), stream));
  ScanTemp<int> t3{s.sort.block_sums, s.sort.counter, num_mix};
  MSMD_CUDA_OK((device_exclusive_scan<int>(LoadInt{s.f3}, SplitCompact{s.r3, offset3, syn3}, n3, t3,
                                           stream)));
  ScanTemp<int> t2{s.sort.block_sums, s.sort.counter, s.total2};
  MSMD_CUDA_OK((device_exclusive_scan<int>(LoadInt{s.f2}, SplitCompact{s.r2, offset2, syn2}, n2, t2,
                                           stream)));
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_compact_unflagged(const int* flags, int n, long long* out_rows, int* count,
                                               void* workspace, size_t workspace_bytes,
                                               msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(n >= 0 && count, "compact_unflagged: bad arguments");
  Workspace ws(workspace, workspace_bytes);
  int* block_sums = ws.take<int>(kScanMaxBlocks);
  unsigned* counter = ws.take<unsigned>(1);
  if (!ws.ok()) {
    set_error("compact_unflagged: workspace too small");
    return MSMD_ERR_WORKSPACE;
  }
  MSMD_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  ScanTemp<int> tmp{block_sums, counter, count};
  MSMD_CUDA_OK((device_exclusive_scan<int>(LoadUnflagged{flags}, StoreRow{out_rows}, n, tmp, stream)));
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_sparse_add_outputs(const int* idx_a, int na, const int* idx_b, int nb,
                                                int batch_size, const int* shape, uint32_t* bits,
                                                int* prefix, int* num_out, void* workspace,
                                                size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(batch_size > 0 && shape[0] > 0 && shape[1] > 0 && shape[2] > 0 && na >= 0 && nb >= 0,
               "sparse_add: bad arguments");
  MSMD_REQUIRE((long long)batch_size * shape[0] * shape[1] * shape[2] < 0x7fffffffLL,
               "sparse_add: grid exceeds the int32 cell index");
  const size_t words = msmd_grid_num_words(batch_size, shape);
  MSMD_CUDA_OK(cudaMemsetAsync(bits, 0, words * sizeof(uint32_t), stream));
  if (na) {
    grid_or_kernel<<<ceil_div(na, 256), 256, 0, stream>>>((const int4*)idx_a, na, batch_size, shape[0],
                                                          shape[1], shape[2], bits);
    MSMD_LAUNCH_OK();
  }
  if (nb) {
    grid_or_kernel<<<ceil_div(nb, 256), 256, 0, stream>>>((const int4*)idx_b, nb, batch_size, shape[0],
                                                          shape[1], shape[2], bits);
    MSMD_LAUNCH_OK();
  }
  Workspace ws(workspace, workspace_bytes);
  int* block_sums = ws.take<int>(kScanMaxBlocks);
  int* total = ws.take<int>(1);
  unsigned* counter = ws.take<unsigned>(1);
  if (!ws.ok()) {
    set_error("sparse_add: workspace too small");
    return MSMD_ERR_WORKSPACE;
  }
  (void)total;
  MSMD_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned), stream));
  ScanTemp<int> tmp{block_sums, counter, num_out};
  MSMD_CUDA_OK((device_exclusive_scan<int>(PopcWord2{bits}, StorePrefix2{prefix}, (int)words, tmp,
                                           stream)));
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_sparse_add_finish(const uint32_t* bits, const int* prefix, int n_out,
                                               const int* idx_a, const float* feat_a, int na,
                                               const int* idx_b, const float* feat_b, int nb, int c,
                                               int batch_size, const int* shape, int* out_indices,
                                               float* out_features, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_out == 0) return MSMD_OK;
  // out_indices / out_features may each be NULL (not both): the coordinates and the features of the sum can be
  // produced by two calls on two streams (csrc/gma.cu: coordinates on the geometry stream)
  MSMD_REQUIRE(c > 0 && (out_indices || out_features), "sparse_add: bad arguments");
  const size_t words = msmd_grid_num_words(batch_size, shape);
  if (out_indices) {
    union_enumerate_kernel<<<ceil_div((long long)words, 256), 256, 0, stream>>>(
        bits, prefix, (int)words, shape[0], shape[1], shape[2], (int4*)out_indices);
    MSMD_LAUNCH_OK();
  }
  if (!out_features) return MSMD_OK;
  MSMD_CUDA_OK(cudaMemsetAsync(out_features, 0, (size_t)n_out * c * sizeof(float), stream));
  if (na) {
    scatter_add_rows_kernel<<<ceil_div((long long)na * c, 256), 256, 0, stream>>>(
        (const int4*)idx_a, feat_a, na, c, batch_size, shape[0], shape[1], shape[2], bits, prefix,
        out_features);
    MSMD_LAUNCH_OK();
  }
  if (nb) {
    scatter_add_rows_kernel<<<ceil_div((long long)nb * c, 256), 256, 0, stream>>>(
        (const int4*)idx_b, feat_b, nb, c, batch_size, shape[0], shape[1], shape[2], bits, prefix,
        out_features);
    MSMD_LAUNCH_OK();
  }
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_lift_gather(const float* img_feat, long long stride_cam,
                                         long long stride_c, long long stride_y, long long stride_x,
                                         int channels, int height, int width, const float* pixels,
                                         const int* cam_ids, const float* points, int point_dims,
                                         int num_points, const float* lidar2img, float downscale,
                                         const float* score_weight, float score_bias, float* out,
                                         msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(channels > 0 && channels <= 128 && point_dims >= 0 && num_points >= 0,
               "lift_gather: channels must be in [1,128]");
  if (num_points == 0) return MSMD_OK;
  lift_gather_kernel<<<ceil_div((long long)num_points * 32, 256), 256, 0, stream>>>(
      img_feat, stride_cam, stride_c, stride_y, stride_x, channels, height, width, pixels, cam_ids,
      points, point_dims, num_points, lidar2img, downscale, score_weight, score_bias, out);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API size_t msmd_rulebook_mask_sort_workspace(int n) {
  Workspace ws((void*)256, ~(size_t)0 >> 1);  // dry run of the carve below
  ws.take<uint32_t>(n > 0 ? n : 1);
  SortWs sw;
  sw.carve(ws, n > 0 ? n : 1);
  return ws.used + 256;
}

extern "C" MSMD_API int msmd_rulebook_mask_sort(const int* pair_fwd, int kvol, int n, int* row_perm,
                                                int* pair_sorted, void* workspace, size_t workspace_bytes,
                                                msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(kvol == 27, "rulebook_mask_sort: the digest is defined for 3x3x3 kernels (kvol = 27)");
  MSMD_REQUIRE(n >= 0, "rulebook_mask_sort: bad size");
  if (n == 0) return MSMD_OK;
  MSMD_REQUIRE(pair_fwd && row_perm && pair_sorted, "rulebook_mask_sort: null pointer");
  Workspace ws(workspace, workspace_bytes);
  uint32_t* keys = ws.take<uint32_t>(n);
  SortWs sw;
  if (!sw.carve(ws, n) || keys == nullptr) {
    set_error("rulebook_mask_sort: workspace too small (%zu bytes given, %zu needed)", workspace_bytes,
              msmd_rulebook_mask_sort_workspace(n));
    return MSMD_ERR_WORKSPACE;
  }
  mask_digest_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(pair_fwd, n, keys);
  MSMD_LAUNCH_OK();
  MSMD_CUDA_OK(radix_sort_pairs(keys, row_perm, n, 16, /*vals_init_iota=*/true, sw, stream));
  pair_permute_kernel<<<ceil_div((long long)kvol * n, 256), 256, 0, stream>>>(pair_fwd, row_perm, kvol, n,
                                                                             pair_sorted);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
