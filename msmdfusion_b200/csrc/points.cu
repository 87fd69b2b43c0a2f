// points.cu -- furthest point sampling, ball query, nearest-3D-voxel search and the
// representative->group assignment used by SparseMultiModalEncoderPaint.fps_NN_fast
// (mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py:276-323).
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace msmd {

// ------------------------------------------------------------------------------------
// Furthest point sampling.
// Reference: one CTA per batch, m-1 serial rounds, each a block-wide arg-max of
// temp[k] = min(temp[k], d(k, last)) (furthest_point_sample_cuda.cu:25-140).  Its result
// depends on the arg-max TIE-BREAK: thread tid = k mod block keeps the first maximal k of
// its strided set (strict >, :69-70) and the shared-memory tree keeps the lower slot on ties
// (__update, :17-23), i.e. among equal distances the winner minimises
//        ( bit_reverse(k mod block),  k / block ).
// Any arg-max under that total order reproduces the reference bit for bit, so the points are
// spread over a thread-block CLUSTER (8 CTAs x 512 threads), coordinates and running
// distances live in registers, and each round reduces a packed 64-bit key
//        (float_bits(dist) << 32) | ~priority
// by warp shuffles, shared memory, and one distributed-shared-memory exchange + cluster
// barrier.  ~2048 rounds stay serial (that is the algorithm) but each round costs a few
// hundred cycles instead of a 50 k-point single-SM sweep.
// ------------------------------------------------------------------------------------
constexpr int kFpsCluster = 8;
constexpr int kFpsThreadsWide = 1024;  // <= 8 points per thread (64 registers): 8*1024*8 = 65536 points
constexpr int kFpsThreadsDeep = 512;   // up to 24 points per thread: 8*512*24 = 98304 points in registers
constexpr int kFpsMaxPerThread = 24;

__device__ __forceinline__ unsigned long long u64max(unsigned long long a, unsigned long long b) {
  return a > b ? a : b;
}

__device__ __forceinline__ unsigned fps_priority(int k, int block, int log2block) {
  const unsigned tid = (unsigned)k & (unsigned)(block - 1);
  const unsigned q = (unsigned)k >> log2block;
  const unsigned rev = log2block ? (__brev(tid) >> (32 - log2block)) : 0u;
  return (rev << 22) | q;  // q < 2^22 is checked on the host
}

// A candidate travels between the CTAs of the cluster as two 16-byte vector stores, each carrying
// the round number as a tag: {key.lo, key.hi, x, tag} {y, z, tag, 0}.  A reader polls its LOCAL
// shared memory until both tags equal the round -- no cluster barrier on the serial chain.
struct __align__(32) FpsCand {
  uint4 h0, h1;
};

__device__ __forceinline__ uint32_t fps_map_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void fps_st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                                  uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ uint4 fps_ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"((uint32_t)__cvta_generic_to_shared(p))
               : "memory");
  return v;
}
// warp-wide max of a 64-bit key with two 32-bit redux instructions (high word first)
__device__ __forceinline__ unsigned long long fps_warp_max(unsigned long long key) {
  const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
  const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
  return ((unsigned long long)mhi << 32) | mlo;
}

template <int PPT, int kFpsThreads>
__global__ void __cluster_dims__(kFpsCluster, 1, 1) __launch_bounds__(kFpsThreads, 1)
fps_cluster_kernel(const float* __restrict__ xyz, int n, int m, int block, int log2block,
                   int* __restrict__ idx) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  extern __shared__ float xyz_s[];  // this CTA's points: slot j*kFpsThreads + tid -> (x,y,z)
  __shared__ unsigned long long warp_best[2][kFpsThreads / 32];  // double-buffered across rounds
  __shared__ FpsCand cand[2][kFpsCluster];                        // slot c is written by CTA c

  const int gtid = rank * kFpsThreads + threadIdx.x;
  const int gthreads = kFpsCluster * kFpsThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float px[PPT], py[PPT], pz[PPT], temp[PPT];
  unsigned prio[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = gtid + j * gthreads;
    if (k < n) {
      px[j] = xyz[3 * k + 0]; py[j] = xyz[3 * k + 1]; pz[j] = xyz[3 * k + 2];
      prio[j] = fps_priority(k, block, log2block);
    } else {
      px[j] = py[j] = pz[j] = 0.f;
      prio[j] = 0xffffffffu;
    }
    float* dst = xyz_s + 3 * (j * kFpsThreads + threadIdx.x);
    dst[0] = px[j]; dst[1] = py[j]; dst[2] = pz[j];
    temp[j] = 1e10f;  // furthest_point_sample.py:28
  }
  if (threadIdx.x < 2 * kFpsCluster) {
    FpsCand z;
    z.h0 = make_uint4(0, 0, 0, 0);
    z.h1 = make_uint4(0, 0, 0, 0);
    (&cand[0][0])[threadIdx.x] = z;  // tag 0 never matches a round number (rounds start at 1)
  }
  // round 0 picks point 0 (furthest_point_sample_cuda.cu:46-47)
  float x1 = __ldg(xyz + 0), y1 = __ldg(xyz + 1), z1 = __ldg(xyz + 2);
  if (gtid == 0) idx[0] = 0;
  cluster.sync();  // every CTA's slots are initialised before anyone publishes into them
  const uint32_t cand_base = (uint32_t)__cvta_generic_to_shared(&cand[0][0]);
  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int k = gtid + j * gthreads;
      if (k < n) {
        const float d = (px[j] - x1) * (px[j] - x1) + (py[j] - y1) * (py[j] - y1) +
                        (pz[j] - z1) * (pz[j] - z1);
        const float d2 = fminf(d, temp[j]);
        temp[j] = d2;
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(~prio[j]);
        best = u64max(best, key);  // "none" is 0; point 0 (priority 0) always has a key > 0
      }
    }
    best = fps_warp_max(best);
    if (lane == 0) warp_best[par][warp] = best;
    __syncthreads();  // the only CTA-wide barrier of the round
    if (warp == 0) {
      unsigned long long b = lane < kFpsThreads / 32 ? warp_best[par][lane] : 0ull;
      b = fps_warp_max(b);
      if (lane < kFpsCluster) {
        // this CTA's best: coordinates from the local shared-memory copy; publish {key, xyz, tag}
        // into slot [rank] of CTA `lane` (distributed shared memory)
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (b != 0ull) {
          const unsigned pr = ~(unsigned)(b & 0xffffffffull);
          const unsigned rev = pr >> 22, q = pr & ((1u << 22) - 1u);
          const unsigned t = log2block ? (__brev(rev) >> (32 - log2block)) : 0u;
          const int k = (int)((q << log2block) | t);
          const int slot = (k / gthreads) * kFpsThreads + (k % gthreads) - (int)rank * kFpsThreads;
          cx = xyz_s[3 * slot + 0]; cy = xyz_s[3 * slot + 1]; cz = xyz_s[3 * slot + 2];
        }
        const uint32_t local = cand_base + (uint32_t)((par * kFpsCluster + rank) * sizeof(FpsCand));
        const uint32_t remote = fps_map_cluster(local, (uint32_t)lane);
        fps_st_cluster_v4(remote, (unsigned)b, (unsigned)(b >> 32), __float_as_uint(cx), (unsigned)r);
        fps_st_cluster_v4(remote + 16, __float_as_uint(cy), __float_as_uint(cz), (unsigned)r, 0u);
      }
    }
    // lanes 0..7 of every warp each wait for one candidate of this round in LOCAL shared memory;
    // the warp then takes the max with two redux instructions and shuffles the winner's xyz
    unsigned long long kc = 0ull;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (lane < kFpsCluster) {
      uint4 h0, h1;
      long long t0 = 0;
      for (;;) {
        h0 = fps_ld_volatile_v4(&cand[par][lane].h0);
        h1 = fps_ld_volatile_v4(&cand[par][lane].h1);
        if (h0.w == (unsigned)r && h1.z == (unsigned)r) break;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000LL) __trap();  // protocol bug: fail, never hang
      }
      kc = ((unsigned long long)h0.y << 32) | h0.x;
      cx = __uint_as_float(h0.z); cy = __uint_as_float(h1.x); cz = __uint_as_float(h1.y);
    }
    __syncwarp();
    const unsigned long long g = fps_warp_max(kc);
    const unsigned winners = __ballot_sync(0xffffffffu, lane < kFpsCluster && kc == g);
    const int src = __ffs(winners) - 1;
    x1 = __shfl_sync(0xffffffffu, cx, src);
    y1 = __shfl_sync(0xffffffffu, cy, src);
    z1 = __shfl_sync(0xffffffffu, cz, src);
    if (gtid == 0) {
      const unsigned p = ~(unsigned)(g & 0xffffffffull);
      const unsigned rev = p >> 22, q = p & ((1u << 22) - 1u);
      const unsigned t = log2block ? (__brev(rev) >> (32 - log2block)) : 0u;
      idx[r] = (int)((q << log2block) | t);
    }
  }
  cluster.sync();  // no CTA may exit while peers can still write into its shared memory
}

// Small point sets (n <= 8192): ONE CTA of 1024 threads, points in registers, a shared-memory copy
// of the coordinates for the winner's lookup.  One __syncthreads per round; every warp reduces the
// 32 per-warp candidates redundantly, so no broadcast step and no inter-SM traffic on the chain.
constexpr int kFpsSingleThreads = 1024;
constexpr int kFpsSingleMaxPerThread = 8;

template <int PPT>
__global__ void __launch_bounds__(kFpsSingleThreads, 1)
fps_single_kernel(const float* __restrict__ xyz, int n, int m, int block, int log2block,
                  int* __restrict__ idx) {
  extern __shared__ float xyz_s[];  // (n, 3)
  __shared__ unsigned long long warp_best[2][kFpsSingleThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;
  float px[PPT], py[PPT], pz[PPT], temp[PPT];
  unsigned prio[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int k = tid + j * nthreads;
    if (k < n) {
      px[j] = xyz[3 * k + 0]; py[j] = xyz[3 * k + 1]; pz[j] = xyz[3 * k + 2];
      xyz_s[3 * k + 0] = px[j]; xyz_s[3 * k + 1] = py[j]; xyz_s[3 * k + 2] = pz[j];
      prio[j] = fps_priority(k, block, log2block);
    } else {
      px[j] = py[j] = pz[j] = 0.f;
      prio[j] = 0xffffffffu;
    }
    temp[j] = 1e10f;
  }
  if (tid == 0) idx[0] = 0;
  __syncthreads();
  float x1 = xyz_s[0], y1 = xyz_s[1], z1 = xyz_s[2];
  for (int r = 1; r < m; ++r) {
    const int par = r & 1;
    unsigned long long best = 0ull;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
      const int k = tid + j * nthreads;
      if (k < n) {
        const float d = (px[j] - x1) * (px[j] - x1) + (py[j] - y1) * (py[j] - y1) +
                        (pz[j] - z1) * (pz[j] - z1);
        const float d2 = fminf(d, temp[j]);
        temp[j] = d2;
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(~prio[j]);
        best = u64max(best, key);
      }
    }
    best = fps_warp_max(best);
    if (lane == 0) warp_best[par][warp] = best;
    __syncthreads();
    const unsigned long long g = fps_warp_max(lane < nwarps ? warp_best[par][lane] : 0ull);
    const unsigned p = ~(unsigned)(g & 0xffffffffull);
    const unsigned rev = p >> 22, q = p & ((1u << 22) - 1u);
    const unsigned t = log2block ? (__brev(rev) >> (32 - log2block)) : 0u;
    const int old = (int)((q << log2block) | t);
    x1 = xyz_s[3 * old + 0]; y1 = xyz_s[3 * old + 1]; z1 = xyz_s[3 * old + 2];
    if (tid == 0) idx[r] = old;
  }
}

// Fallback for point sets that do not fit the register-resident cluster kernel: one CTA,
// temp[] in global memory, same tie-break rule.
__global__ void __launch_bounds__(1024, 1)
fps_global_kernel(const float* __restrict__ xyz, int n, int m, int block, int log2block,
                  float* __restrict__ temp, int* __restrict__ idx) {
  __shared__ unsigned long long warp_best[32];
  __shared__ int s_old;
  for (int k = threadIdx.x; k < n; k += blockDim.x) temp[k] = 1e10f;
  if (threadIdx.x == 0) idx[0] = 0;
  __syncthreads();
  int old = 0;
  for (int r = 1; r < m; ++r) {
    const float x1 = xyz[3 * old], y1 = xyz[3 * old + 1], z1 = xyz[3 * old + 2];
    unsigned long long best = 0ull;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const float x2 = xyz[3 * k], y2 = xyz[3 * k + 1], z2 = xyz[3 * k + 2];
      const float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
      const float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) |
                                     (unsigned long long)(~fps_priority(k, block, log2block));
      best = u64max(best, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = u64max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) warp_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long b = threadIdx.x < (blockDim.x >> 5) ? warp_best[threadIdx.x] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = u64max(b, __shfl_xor_sync(0xffffffffu, b, o));
      if (threadIdx.x == 0) {
        const unsigned p = ~(unsigned)(b & 0xffffffffull);
        const unsigned rev = p >> 22, q = p & ((1u << 22) - 1u);
        const unsigned tid = log2block ? (__brev(rev) >> (32 - log2block)) : 0u;
        s_old = (int)((q << log2block) | tid);
        idx[r] = s_old;
      }
    }
    __syncthreads();
    old = s_old;
  }
}

// ------------------------------------------------------------------------------------
// Ball query (ball_query_cuda.cu:11-55): first `nsample` candidate indices (ascending) with
// d2 == 0 || min_r2 <= d2 < max_r2, padded with the first hit; rows with no hit keep the
// caller's zero initialisation (ball_query.py:35).  One warp per centre: 32 candidates per
// step, ballot, ordered append, early exit.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ball_query_kernel(const float* __restrict__ xyz, int n, const float* __restrict__ centers, int m,
                  float min_r2, float max_r2, int nsample, int* __restrict__ idx) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= m) return;
  const float nx = centers[3 * c], ny = centers[3 * c + 1], nz = centers[3 * c + 2];
  int* row = idx + (size_t)c * nsample;
  int cnt = 0;
  int first = -1;
  // kBqUnroll groups of 32 candidates are loaded before the first ballot, so that many L2 round trips
  // overlap (the scan is latency bound: one dependent ballot per step).  The groups are then appended
  // in order; a group that starts after the row is full writes nothing (pos >= nsample).
  constexpr int kBqUnroll = 4;
  for (int base = 0; base < n && cnt < nsample; base += 32 * kBqUnroll) {
    bool hit[kBqUnroll];
#pragma unroll
    for (int u = 0; u < kBqUnroll; ++u) {
      const int k = base + 32 * u + lane;
      hit[u] = false;
      if (k < n) {
        const float x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
        const float d2 = (nx - x) * (nx - x) + (ny - y) * (ny - y) + (nz - z) * (nz - z);
        hit[u] = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
      }
    }
#pragma unroll
    for (int u = 0; u < kBqUnroll; ++u) {
      const unsigned bal = __ballot_sync(0xffffffffu, hit[u]);
      if (bal) {
        if (first < 0) first = base + 32 * u + __ffs(bal) - 1;
        const int pos = cnt + __popc(bal & ((1u << lane) - 1u));
        if (hit[u] && pos < nsample) row[pos] = base + 32 * u + lane;
        cnt += __popc(bal);
      }
    }
  }
  if (first >= 0) {
    if (cnt > nsample) cnt = nsample;
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;
  }
}

// ------------------------------------------------------------------------------------
// Nearest key per query on integer voxel coordinates
// (sparse_multimodal_encoder_painting.py:289-291 / :302-305): dist = ||q - key||_2 in fp32,
// (val, idx) = min over keys with the FIRST minimal index winning.  Squared distances of
// voxel coordinates are exact integers (< 2^23), so the arg-min is taken on
// (d2 << 32 | index) and val = sqrtf(d2).  One CTA per query, keys streamed coalesced.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nn_search_kernel(const int* __restrict__ query, int q_stride, int nq, const int* __restrict__ key,
                 int k_stride, int nk, float* __restrict__ val, int* __restrict__ idx) {
  __shared__ unsigned long long wbest[8];
  const int q = blockIdx.x;
  if (q >= nq) return;
  const int qz = query[(size_t)q * q_stride + 0], qy = query[(size_t)q * q_stride + 1],
            qx = query[(size_t)q * q_stride + 2];
  unsigned long long best = ~0ull;
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    const int* kp = key + (size_t)k * k_stride;
    const long long dz = kp[0] - qz, dy = kp[1] - qy, dx = kp[2] - qx;
    const unsigned long long d2 = (unsigned long long)(dz * dz + dy * dy + dx * dx);
    const unsigned long long cand = (d2 << 32) | (unsigned)k;
    best = cand < best ? cand : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0) wbest[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best = wbest[w] < best ? wbest[w] : best;
    if (nk > 0) {
      val[q] = sqrtf((float)(best >> 32));
      idx[q] = (int)(best & 0xffffffffull);
    } else {
      val[q] = INFINITY;
      idx[q] = 0;
    }
  }
}

// Resolve `query_NN_key_idx[group] = nn` (painting.py:311-321).  The reference scatter has
// duplicate indices (winner undefined); here the LAST (representative, slot) pair in
// row-major order wins: pass 1 records the largest flat position per only-2D voxel, pass 2
// reads that representative's nearest key.
__global__ void __launch_bounds__(256)
group_winner_kernel(const int* __restrict__ group, int m, int nsample, const float* __restrict__ val,
                    float thresh, int* __restrict__ winner) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m * nsample) return;
  const int rep = t / nsample;
  if (!(val[rep] < thresh)) return;
  atomicMax(&winner[group[t]], t);
}

__global__ void __launch_bounds__(256)
group_assign_kernel(const int* __restrict__ winner, int nq, int nsample,
                    const int* __restrict__ nn_idx, int base, long long* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int w = winner[q];
  out[q] = (w < 0) ? -1ll : (long long)(nn_idx[w / nsample] + base);
}

__global__ void __launch_bounds__(256)
direct_assign_kernel(const float* __restrict__ val, const int* __restrict__ nn_idx, int nq,
                     float thresh, int base, long long* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  out[q] = (val[q] < thresh) ? (long long)(nn_idx[q] + base) : -1ll;
}

static int fps_block_size(int n, int* log2block) {
  // opt_n_threads (furthest_point_sample_cuda.cu:11-15): truncating log2 computed in double
  int pow_2 = (int)(log((double)n) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  int l = 0;
  while ((1 << l) < t) ++l;
  *log2block = l;
  return t;
}

}  // namespace msmd

using namespace msmd;

static int g_fps_threads = 0;
// A/B switch of the cluster FPS kernel's CTA width: 0 (default heuristic) | 256 | 512 | 1024
extern "C" MSMD_API int msmd_fps_set_threads(int threads) {
  MSMD_REQUIRE(threads == 0 || threads == 256 || threads == 512 || threads == 1024, "fps_set_threads: 0, 256, 512 or 1024");
  g_fps_threads = threads;
  return MSMD_OK;
}

extern "C" MSMD_API size_t msmd_fps_workspace(int n) { return (size_t)(n > 0 ? n : 1) * sizeof(float) + 256; }

extern "C" MSMD_API int msmd_fps(const float* xyz, int n, int m, int* idx, void* workspace,
                                 size_t workspace_bytes, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(xyz && idx && n > 0 && m > 0, "fps: bad arguments");
  int log2block = 0;
  const int block = fps_block_size(n, &log2block);
  MSMD_REQUIRE((n >> log2block) < (1 << 22), "fps: too many points");
  const long long cap_wide = (long long)kFpsCluster * kFpsThreadsWide, cap_deep = (long long)kFpsCluster * kFpsThreadsDeep;
  if (n <= kFpsSingleThreads * kFpsSingleMaxPerThread) {
    // the round is instruction-issue bound (every warp repeats the block-level reduction), so
    // use as few warps as keep <= 8 points per thread
    int threads = ceil_div(ceil_div(n, kFpsSingleMaxPerThread), 32) * 32;
    if (threads < 128) threads = 128;
    if (threads > kFpsSingleThreads) threads = kFpsSingleThreads;
    const size_t smem = (size_t)n * 3 * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
      MSMD_CUDA_OK(cudaFuncSetAttribute(fps_single_kernel<kFpsSingleMaxPerThread>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_done = true;
    }
    fps_single_kernel<kFpsSingleMaxPerThread><<<1, threads, smem, stream>>>(xyz, n, m, block, log2block, idx);
    MSMD_LAUNCH_OK();
  } else if (n <= cap_deep * kFpsMaxPerThread) {
    // the per-round cost is (points per thread) x ~13 instructions + a fixed reduction / exchange chain, and the
    // guarded iterations of a template that is deeper than needed are not free: 1024-thread CTAs with an exact
    // per-thread depth whenever the points fit 8 per thread, the 512-thread deep variants above that
#define MSMD_FPS(P, T)                                                                          \
  do {                                                                                          \
    const size_t smem = (size_t)(P) * (T) * 3 * sizeof(float);                                  \
    static bool attr_done = false; /* dynamic + 640 B static must fit: raise the limit once */  \
    if (!attr_done) {                                                                           \
      MSMD_CUDA_OK(cudaFuncSetAttribute(fps_cluster_kernel<P, T>,                               \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      attr_done = true;                                                                         \
    }                                                                                           \
    fps_cluster_kernel<P, T><<<kFpsCluster, T, smem, stream>>>(xyz, n, m, block, log2block, idx); \
  } while (0)
    // A/B switch (msmd_fps_set_threads): 0 = the default below; 256 / 512 / 1024 = force that CTA width when the points fit
    const long long cap_thin = (long long)kFpsCluster * 256;
    int width = g_fps_threads;
    // measured on the LC scene (profiles/r02k_lc_timeline_fps*.txt): 46.8k points 2.36 ms at 512 threads vs 2.68 at
    // 1024 and 2.95 at 256; 35k points 1.82 vs 1.96; 24k points equal -> the wide CTA only up to 3 points per thread
    if (width == 0) width = (n <= cap_wide * 3) ? 1024 : 512;
    if (width == 256 && n > cap_thin * kFpsMaxPerThread) width = 512;
    if (width == 1024 && n > cap_wide * 8) width = 512;
    if (width == 1024) {
      const int ppt = ceil_div(n, cap_wide);
      if (ppt <= 1) MSMD_FPS(1, kFpsThreadsWide);
      else if (ppt <= 2) MSMD_FPS(2, kFpsThreadsWide);
      else if (ppt <= 3) MSMD_FPS(3, kFpsThreadsWide);
      else if (ppt <= 4) MSMD_FPS(4, kFpsThreadsWide);
      else if (ppt <= 6) MSMD_FPS(6, kFpsThreadsWide);
      else MSMD_FPS(8, kFpsThreadsWide);
    } else if (width == 256) {
      const int ppt = ceil_div(n, cap_thin);
      if (ppt <= 4) MSMD_FPS(4, 256);
      else if (ppt <= 8) MSMD_FPS(8, 256);
      else if (ppt <= 12) MSMD_FPS(12, 256);
      else if (ppt <= 16) MSMD_FPS(16, 256);
      else MSMD_FPS(kFpsMaxPerThread, 256);
    } else {
      const int ppt = ceil_div(n, cap_deep);
      if (ppt <= 4) MSMD_FPS(4, kFpsThreadsDeep);
      else if (ppt <= 8) MSMD_FPS(8, kFpsThreadsDeep);
      else if (ppt <= 12) MSMD_FPS(12, kFpsThreadsDeep);
      else if (ppt <= 16) MSMD_FPS(16, kFpsThreadsDeep);
      else MSMD_FPS(kFpsMaxPerThread, kFpsThreadsDeep);
    }
#undef MSMD_FPS
    MSMD_LAUNCH_OK();
  } else {
    if (!workspace || workspace_bytes < msmd_fps_workspace(n)) {
      set_error("fps: workspace too small");
      return MSMD_ERR_WORKSPACE;
    }
    fps_global_kernel<<<1, 1024, 0, stream>>>(xyz, n, m, block, log2block, (float*)workspace, idx);
    MSMD_LAUNCH_OK();
  }
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_ball_query(const float* xyz, int n, const float* centers, int m,
                                        float min_radius, float max_radius, int nsample, int* idx,
                                        msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(idx && nsample > 0 && n >= 0 && m >= 0, "ball_query: bad arguments");
  if (m == 0) return MSMD_OK;
  MSMD_CUDA_OK(cudaMemsetAsync(idx, 0, (size_t)m * nsample * sizeof(int), stream));
  if (n == 0) return MSMD_OK;
  ball_query_kernel<<<ceil_div((long long)m * 32, 256), 256, 0, stream>>>(
      xyz, n, centers, m, min_radius * min_radius, max_radius * max_radius, nsample, idx);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_nn_search(const int* query, int query_stride, int nq, const int* key,
                                       int key_stride, int nk, float* val, int* idx,
                                       msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(nq >= 0 && nk >= 0 && query_stride >= 3 && key_stride >= 3, "nn_search: bad arguments");
  if (nq == 0) return MSMD_OK;
  nn_search_kernel<<<nq, 256, 0, stream>>>(query, query_stride, nq, key, key_stride, nk, val, idx);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_group_assign(const int* group, int m, int nsample, const float* val,
                                          const int* nn_idx, float dist_thresh, int nq, int base,
                                          int* winner_scratch, long long* out,
                                          msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(nq >= 0 && m >= 0, "group_assign: bad arguments");
  if (nq == 0) return MSMD_OK;
  if (group == nullptr) {  // Q <= fps_num: every query is its own representative
    direct_assign_kernel<<<ceil_div(nq, 256), 256, 0, stream>>>(val, nn_idx, nq, dist_thresh, base, out);
    MSMD_LAUNCH_OK();
    return MSMD_OK;
  }
  MSMD_CUDA_OK(cudaMemsetAsync(winner_scratch, 0xFF, (size_t)nq * sizeof(int), stream));
  if (m > 0) {
    group_winner_kernel<<<ceil_div((long long)m * nsample, 256), 256, 0, stream>>>(
        group, m, nsample, val, dist_thresh, winner_scratch);
    MSMD_LAUNCH_OK();
  }
  group_assign_kernel<<<ceil_div(nq, 256), 256, 0, stream>>>(winner_scratch, nq, nsample, nn_idx, base, out);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
