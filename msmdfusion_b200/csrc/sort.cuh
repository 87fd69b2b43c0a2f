// sort.cuh -- stable LSD radix sort of (uint32 key, int32 value) pairs, 8 bits per pass.
//
// Used where the reference sorts (torch.sort in voxel_modality_split, MSMDFusion.py:274-275)
// or where a deterministic order is needed.  Stability makes the result a well-defined
// function of the input (ties keep input order), unlike the reference's unstable sort.
//
// Per pass: (1) per-CTA digit histogram -> hist[digit][cta]; (2) device-wide exclusive scan of
// that digit-major matrix; (3) stable scatter: inside a CTA the tile is walked in chunks of
// 256 keys, a key's rank among equal digits = keys of earlier chunks + keys of lower warps
// (shared counters) + lower lanes of its own warp (__match_any_sync).
#pragma once
#include "scan.cuh"

namespace msmd {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;  // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kSortRadix = 256;

static __global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift, int* __restrict__ hist,
                 int nblocks) {
  __shared__ int h[kSortRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
  for (int j = 0; j < kSortItems; ++j) {
    const int i = base + j * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1);
  }
  __syncthreads();
  hist[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

struct LoadInt {
  const int* p;
  __device__ int operator()(int i) const { return p[i]; }
};
struct StoreInt {
  int* p;
  __device__ void operator()(int i, int ex, int) const { p[i] = ex; }
};

static __global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint32_t* __restrict__ keys_in, const int* __restrict__ vals_in, int n,
                    int shift, const int* __restrict__ offsets, int nblocks,
                    uint32_t* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int base[kSortRadix];                 // next free global slot per digit
  __shared__ int wcnt[kSortThreads / 32][kSortRadix];  // per-warp digit counts of this chunk
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  base[tid] = offsets[tid * nblocks + blockIdx.x];
#pragma unroll
  for (int j = 0; j < kSortThreads / 32; ++j) wcnt[j][tid] = 0;
  __syncthreads();
  const int tile = blockIdx.x * kSortTile;
  for (int j = 0; j < kSortItems; ++j) {
    const int i = tile + j * kSortThreads + tid;
    const bool valid = i < n;
    uint32_t key = 0;
    int val = 0;
    int digit = -1;
    if (valid) {
      key = keys_in[i];
      val = vals_in ? vals_in[i] : i;
      digit = (key >> shift) & 0xFF;
    }
    // lanes holding the same digit (invalid lanes share digit -1 and are ignored)
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) wcnt[w][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      int off = base[digit] + rank_in_warp;
      for (int ww = 0; ww < w; ++ww) off += wcnt[ww][digit];
      keys_out[off] = key;
      vals_out[off] = val;
    }
    __syncthreads();
    {
      int add = 0;
#pragma unroll
      for (int ww = 0; ww < kSortThreads / 32; ++ww) {
        add += wcnt[ww][tid];
        wcnt[ww][tid] = 0;
      }
      base[tid] += add;
    }
    __syncthreads();
  }
}

static inline int sort_num_blocks(int n) { return ceil_div(n > 0 ? n : 1, kSortTile); }

// scratch: hist ints [256 * nblocks] + ping-pong key/value buffers [n] each + scan temp
struct SortWs {
  int* hist;
  uint32_t* keys_tmp;
  int* vals_tmp;
  int* block_sums;
  int* total;
  unsigned* counter;
  bool carve(Workspace& ws, int n) {
    const int nb = sort_num_blocks(n);
    hist = ws.take<int>((size_t)kSortRadix * nb);
    keys_tmp = ws.take<uint32_t>(n > 0 ? n : 1);
    vals_tmp = ws.take<int>(n > 0 ? n : 1);
    block_sums = ws.take<int>(kScanMaxBlocks);
    total = ws.take<int>(1);
    counter = ws.take<unsigned>(1);
    return ws.ok();
  }
};

// Sorts `keys` (in place) ascending and produces `vals` (in place; if vals_init_iota the
// input values are 0..n-1).  `key_bits` = number of significant key bits (<= 32).
// The number of passes is rounded up to an even count so the result lands in keys/vals.
static inline cudaError_t radix_sort_pairs(uint32_t* keys, int* vals, int n, int key_bits,
                                           bool vals_init_iota, SortWs& s, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  if (passes & 1) ++passes;
  const int nb = sort_num_blocks(n);
  cudaError_t e = cudaMemsetAsync(s.counter, 0, sizeof(unsigned), stream);
  if (e != cudaSuccess) return e;
  uint32_t* kin = keys;
  int* vin = vals;
  uint32_t* kout = s.keys_tmp;
  int* vout = s.vals_tmp;
  for (int p = 0; p < passes; ++p) {
    const int shift = 8 * p;
    sort_hist_kernel<<<nb, kSortThreads, 0, stream>>>(kin, n, shift, s.hist, nb);
    count_launch(1);
    ScanTemp<int> tmp{s.block_sums, s.counter, s.total};
    e = device_exclusive_scan<int>(LoadInt{s.hist}, StoreInt{s.hist}, kSortRadix * nb, tmp, stream);
    if (e != cudaSuccess) return e;
    sort_scatter_kernel<<<nb, kSortThreads, 0, stream>>>(
        kin, (p == 0 && vals_init_iota) ? nullptr : vin, n, shift, s.hist, nb, kout, vout);
    count_launch(1);
    uint32_t* tk = kin; kin = kout; kout = tk;
    int* tv = vin; vin = vout; vout = tv;
  }
  return cudaGetLastError();
}

}  // namespace msmd
