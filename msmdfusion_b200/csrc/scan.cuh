// scan.cuh -- device-wide exclusive scan over a functor, two launches, no spin-waits.
//
//   launch 1 (reduce): each CTA sums a contiguous chunk; the LAST CTA to finish (ticket
//                      counter + __threadfence) scans the <=512 CTA sums in place and
//                      publishes the grand total.
//   launch 2 (apply) : each CTA re-scans its chunk from its CTA offset and hands
//                      (index, exclusive prefix, value) to the output functor.
//
// The input functor is evaluated twice; all users read L2-resident arrays (bitmaps,
// per-point flags), so the second read never reaches HBM on B200's 126 MB L2.
#pragma once
#include "common.cuh"

namespace msmd {

constexpr int kScanThreads = 512;
constexpr int kScanMaxBlocks = 512;  // one CTA sum per thread of the finishing CTA

template <typename T>
struct ScanTemp {
  T* block_sums;      // [kScanMaxBlocks]
  unsigned* counter;  // zero on entry; reset to zero by the finishing CTA
  T* total;           // device scalar: sum of all values
};

template <typename T, typename F>
__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(F f, int n, int chunk, ScanTemp<T> tmp) {
  __shared__ T sm[33];
  __shared__ int is_last;
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);
  T s = T(0);
  for (int i = beg + (int)threadIdx.x; i < end; i += blockDim.x) s += f(i);
  T tot;
  block_exclusive_scan(s, tot, sm);
  if (threadIdx.x == 0) {
    tmp.block_sums[blockIdx.x] = tot;
    __threadfence();
    const unsigned ticket = atomicAdd(tmp.counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const volatile T* bs = tmp.block_sums;
    T v = (threadIdx.x < gridDim.x) ? bs[threadIdx.x] : T(0);
    T tot2;
    T ex = block_exclusive_scan(v, tot2, sm);
    if (threadIdx.x < gridDim.x) tmp.block_sums[threadIdx.x] = ex;
    if (threadIdx.x == 0) {
      *tmp.total = tot2;
      *tmp.counter = 0u;
    }
  }
}

template <typename T, typename F, typename O>
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(F f, O o, int n, int chunk, ScanTemp<T> tmp) {
  __shared__ T sm[33];
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);
  T run = tmp.block_sums[blockIdx.x];
  for (int base = beg; base < end; base += blockDim.x) {
    const int i = base + (int)threadIdx.x;
    const T v = (i < end) ? f(i) : T(0);
    T tot;
    const T ex = block_exclusive_scan(v, tot, sm);
    if (i < end) o(i, run + ex, v);
    run += tot;
  }
}

template <typename T, typename F, typename O>
static inline cudaError_t device_exclusive_scan(F f, O o, int n, ScanTemp<T> tmp,
                                                cudaStream_t stream) {
  if (n <= 0) return cudaMemsetAsync(tmp.total, 0, sizeof(T), stream);
  int blocks = ceil_div(n, kScanThreads);
  if (blocks > kScanMaxBlocks) blocks = kScanMaxBlocks;
  int chunk = ceil_div(n, blocks);
  chunk = ceil_div(chunk, kScanThreads) * kScanThreads;
  blocks = ceil_div(n, chunk);
  scan_reduce_kernel<T, F><<<blocks, kScanThreads, 0, stream>>>(f, n, chunk, tmp);
  scan_apply_kernel<T, F, O><<<blocks, kScanThreads, 0, stream>>>(f, o, n, chunk, tmp);
  count_launch(2);
  return cudaGetLastError();
}

}  // namespace msmd
