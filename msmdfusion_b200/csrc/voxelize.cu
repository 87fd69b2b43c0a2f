// voxelize.cu -- hash-based hard_voxelize with the HardSimpleVFE mean fused in.
//
// Reference semantics (mmdet3d/ops/voxel/src/voxelization_cpu.cpp:43-99 and the CUDA
// path src/voxelization_cuda.cu:105-180): voxels are numbered in order of FIRST
// APPEARANCE in the input; each voxel keeps its first `max_points` points in input order;
// when a new voxel would exceed `max_voxels` ALL further processing stops (`break`).
//
// The reference GPU path gets this order from an O(N^2) duplicate scan plus a <<<1,1>>>
// serial pass and four device synchronisations.  Here:
//   1. insert   : per point, IEEE fp32 floor((p-min)/size) -> linear key -> open-addressing
//                 hash (atomicCAS); atomicMin records the voxel's first point, atomicAdd its
//                 population.
//   2. scan     : a point is a "voxel head" iff it is its voxel's first point.  An exclusive
//                 scan of the head flags over POINT order is exactly the first-appearance
//                 rank; the same scan (upper 32 bits) allocates each voxel's bucket.  The head
//                 whose rank == max_voxels defines the cut-off point index (the `break`).
//   3. bucket   : points before the cut-off drop their index into their voxel's bucket.
//   4. gather   : one warp per voxel extracts the <= max_points smallest point indices in
//                 ascending order (== input order), copies the rows, and accumulates the
//                 mean in slot order.
// No host synchronisation; voxel_num stays on the device.
#include "scan.cuh"

namespace msmd {

struct VoxGeom {
  float vx, vy, vz;
  float minx, miny, minz;
  int gx, gy, gz;
};

__device__ __forceinline__ unsigned hash_u32(unsigned x) {
  x ^= x >> 16; x *= 0x85ebca6bu;
  x ^= x >> 13; x *= 0xc2b2ae35u;
  x ^= x >> 16;
  return x;
}

// floor((p - min) / size) exactly as the reference computes it: IEEE fp32 subtract and
// true division (voxelization_cpu.cpp:23); NaN / +-inf / out-of-range all fail.
__device__ __forceinline__ bool voxel_coord(float p, float mn, float sz, int grid, int& c) {
  const float f = floorf(__fdiv_rn(__fsub_rn(p, mn), sz));
  const bool ok = (f >= 0.0f) && (f < (float)grid);
  c = ok ? (int)f : -1;
  return ok;
}

__global__ void __launch_bounds__(256)
vox_insert_kernel(const float* __restrict__ points, int n, int C, VoxGeom g,
                  int* __restrict__ hkeys, int* __restrict__ hfirst, int* __restrict__ hcount,
                  unsigned mask, int* __restrict__ pslot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = points + (size_t)i * C;
  int cx, cy, cz;
  const bool ok = voxel_coord(p[0], g.minx, g.vx, g.gx, cx) &&
                  voxel_coord(p[1], g.miny, g.vy, g.gy, cy) &&
                  voxel_coord(p[2], g.minz, g.vz, g.gz, cz);
  if (!ok) {
    pslot[i] = -1;
    return;
  }
  const int key = (cz * g.gy + cy) * g.gx + cx;
  unsigned slot = hash_u32((unsigned)key) & mask;
  while (true) {
    const int prev = atomicCAS(&hkeys[slot], -1, key);
    if (prev == -1 || prev == key) break;
    slot = (slot + 1) & mask;
  }
  atomicMin(&hfirst[slot], i);
  atomicAdd(&hcount[slot], 1);
  pslot[i] = (int)slot;
}

struct VoxHeadFlag {
  const int* pslot;
  const int* hfirst;
  const int* hcount;
  __device__ unsigned long long operator()(int i) const {
    const int s = pslot[i];
    if (s < 0 || hfirst[s] != i) return 0ull;
    return 1ull | ((unsigned long long)(unsigned)hcount[s] << 32);
  }
};

struct VoxAssign {
  const int* pslot;
  const int* hkeys;
  int* hoff;
  int* vslot;
  int* coors;
  int ncol;
  int batch_idx;
  int max_voxels;
  int gx, gy;
  int* cutoff;
  __device__ void operator()(int i, unsigned long long ex, unsigned long long v) const {
    if (v == 0ull) return;
    const int rank = (int)(unsigned)(ex & 0xffffffffull);
    const int off = (int)(unsigned)(ex >> 32);
    const int s = pslot[i];
    hoff[s] = off;
    if (rank < max_voxels) {
      vslot[rank] = s;
      const int key = hkeys[s];
      const int x = key % gx;
      const int y = (key / gx) % gy;
      const int z = key / (gx * gy);
      int* c = coors + (size_t)rank * ncol;
      if (ncol == 4) {
        c[0] = batch_idx; c[1] = z; c[2] = y; c[3] = x;
      } else {
        c[0] = z; c[1] = y; c[2] = x;
      }
    } else if (rank == max_voxels) {
      *cutoff = i;  // the point at which the reference loop `break`s
    }
  }
};

__global__ void __launch_bounds__(256)
vox_bucket_kernel(int n, const int* __restrict__ pslot, const int* __restrict__ hoff,
                  int* __restrict__ hcursor, int* __restrict__ bucket,
                  const int* __restrict__ cutoff, const unsigned long long* __restrict__ total,
                  int max_voxels, int* __restrict__ voxel_num) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    const int nv = (int)(unsigned)(*total & 0xffffffffull);
    *voxel_num = nv < max_voxels ? nv : max_voxels;
  }
  if (i >= n) return;
  const int s = pslot[i];
  if (s < 0 || i >= *cutoff) return;
  const int pos = atomicAdd(&hcursor[s], 1);
  bucket[hoff[s] + pos] = i;
}

constexpr int kVoxMaxChunks = 8;  // supports num_features <= 256

__global__ void __launch_bounds__(256)
vox_gather_kernel(const float* __restrict__ points, int C, int max_points, int max_voxels,
                  const unsigned long long* __restrict__ total, const int* __restrict__ vslot,
                  const int* __restrict__ hoff, const int* __restrict__ hcursor,
                  const int* __restrict__ bucket, float* __restrict__ voxels,
                  int* __restrict__ num_out, float* __restrict__ mean, int F) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  int nv = (int)(unsigned)(*total & 0xffffffffull);
  nv = nv < max_voxels ? nv : max_voxels;
  if (w >= nv) return;
  const int s = vslot[w];
  const int cnt = hcursor[s];
  const int* b = bucket + hoff[s];
  const int num = cnt < max_points ? cnt : max_points;

  float sum[kVoxMaxChunks];
#pragma unroll
  for (int j = 0; j < kVoxMaxChunks; ++j) sum[j] = 0.f;

  int prev = -1;
  for (int r = 0; r < num; ++r) {
    // next smallest point index > prev
    int m = 0x7fffffff;
    for (int j = lane; j < cnt; j += 32) {
      const int e = b[j];
      if (e > prev && e < m) m = e;
    }
    m = warp_min(m);
    prev = m;
    const float* src = points + (size_t)m * C;
    float* dst = voxels ? voxels + ((size_t)w * max_points + r) * C : nullptr;
#pragma unroll
    for (int j = 0; j < kVoxMaxChunks; ++j) {
      const int c = lane + 32 * j;
      if (c < C) {
        const float v = src[c];
        if (dst) dst[c] = v;
        sum[j] += v;  // slot order, like features.sum(dim=1)
      }
    }
  }
  if (voxels) {
    float* z = voxels + ((size_t)w * max_points + num) * C;
    const int nz = (max_points - num) * C;
    for (int j = lane; j < nz; j += 32) z[j] = 0.f;
  }
  if (lane == 0) num_out[w] = num;
  if (mean) {
    const float denom = (float)num;
#pragma unroll
    for (int j = 0; j < kVoxMaxChunks; ++j) {
      const int c = lane + 32 * j;
      if (c < F) mean[(size_t)w * F + c] = __fdiv_rn(sum[j], denom);
    }
  }
}

static unsigned hash_capacity(int n) {
  unsigned cap = 1024;
  while (cap < 2u * (unsigned)n) cap <<= 1;
  return cap;
}

struct VoxWorkspace {
  int *hkeys, *hfirst, *hcount, *hcursor, *hoff, *pslot, *vslot, *bucket, *cutoff;
  unsigned long long* block_sums;
  unsigned long long* total;
  unsigned* counter;
  unsigned cap;
  bool carve(Workspace& ws, int n) {
    cap = hash_capacity(n);
    // [hkeys | hfirst] and [hcount | hcursor | counter] are adjacent so that two memsets
    // initialise the table.
    hkeys = ws.take<int>(cap);
    hfirst = ws.take<int>(cap);
    hcount = ws.take<int>(cap);
    hcursor = ws.take<int>(cap);
    hoff = ws.take<int>(cap);
    pslot = ws.take<int>(n);
    vslot = ws.take<int>(n);
    bucket = ws.take<int>(n);
    cutoff = ws.take<int>(1);
    block_sums = ws.take<unsigned long long>(kScanMaxBlocks);
    total = ws.take<unsigned long long>(1);
    counter = ws.take<unsigned>(1);
    return ws.ok();
  }
};

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API size_t msmd_hard_voxelize_workspace(int num_points) {
  if (num_points < 1) num_points = 1;
  Workspace ws((void*)256, ~(size_t)0 >> 1);  // dry run: only offsets matter
  VoxWorkspace v;
  v.carve(ws, num_points);
  return ws.used + 256;
}

extern "C" MSMD_API int msmd_hard_voxelize(const float* points, int num_points, int num_features,
                                  const float* voxel_size, const float* coors_range,
                                  int max_points, int max_voxels, float* voxels, int* coors,
                                  int coors_ncol, int batch_idx, int* num_points_per_voxel,
                                  float* mean, int mean_features, int* voxel_num,
                                  void* workspace, size_t workspace_bytes,
                                  msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(num_points >= 0 && num_features >= 3, "hard_voxelize: need num_features >= 3");
  MSMD_REQUIRE(num_features <= 32 * kVoxMaxChunks, "hard_voxelize: num_features > %d unsupported",
               32 * kVoxMaxChunks);
  MSMD_REQUIRE(coors_ncol == 3 || coors_ncol == 4, "hard_voxelize: coors_ncol must be 3 or 4");
  MSMD_REQUIRE(max_points >= 1, "hard_voxelize: max_points must be >= 1 (dynamic voxelization is "
                                "out of scope)");
  MSMD_REQUIRE(voxel_num && coors && num_points_per_voxel, "hard_voxelize: null output");
  MSMD_REQUIRE(mean == nullptr || (mean_features >= 1 && mean_features <= num_features),
               "hard_voxelize: bad mean_features");
  if (max_voxels < 0) max_voxels = 0x7fffffff;  // reference: -1 == unlimited
  if (num_points == 0) {
    MSMD_CUDA_OK(cudaMemsetAsync(voxel_num, 0, sizeof(int), stream));
    return MSMD_OK;
  }
  VoxGeom g;
  g.vx = voxel_size[0]; g.vy = voxel_size[1]; g.vz = voxel_size[2];
  g.minx = coors_range[0]; g.miny = coors_range[1]; g.minz = coors_range[2];
  // grid_size = round((max - min) / size) in fp32 (voxelization_cpu.cpp:121-124)
  g.gx = (int)roundf((coors_range[3] - coors_range[0]) / voxel_size[0]);
  g.gy = (int)roundf((coors_range[4] - coors_range[1]) / voxel_size[1]);
  g.gz = (int)roundf((coors_range[5] - coors_range[2]) / voxel_size[2]);
  MSMD_REQUIRE(g.gx > 0 && g.gy > 0 && g.gz > 0, "hard_voxelize: empty grid");
  MSMD_REQUIRE((long long)g.gx * g.gy * g.gz < 0x7fffffffLL,
               "hard_voxelize: grid volume exceeds int32 keys");

  Workspace ws(workspace, workspace_bytes);
  VoxWorkspace v;
  if (!v.carve(ws, num_points)) {
    set_error("hard_voxelize: workspace too small (%zu < %zu)", workspace_bytes,
              msmd_hard_voxelize_workspace(num_points));
    return MSMD_ERR_WORKSPACE;
  }
  const size_t tbl = (size_t)v.cap * sizeof(int);
  MSMD_CUDA_OK(cudaMemsetAsync(v.hkeys, 0xFF, tbl, stream));   // -1: empty
  MSMD_CUDA_OK(cudaMemsetAsync(v.hfirst, 0x7F, tbl, stream));  // 0x7f7f7f7f > any index
  MSMD_CUDA_OK(cudaMemsetAsync(v.hcount, 0, tbl, stream));
  MSMD_CUDA_OK(cudaMemsetAsync(v.hcursor, 0, tbl, stream));
  MSMD_CUDA_OK(cudaMemsetAsync(v.cutoff, 0x7F, sizeof(int), stream));
  MSMD_CUDA_OK(cudaMemsetAsync(v.counter, 0, sizeof(unsigned), stream));

  const int tpb = 256;
  vox_insert_kernel<<<ceil_div(num_points, tpb), tpb, 0, stream>>>(
      points, num_points, num_features, g, v.hkeys, v.hfirst, v.hcount, v.cap - 1, v.pslot);
  MSMD_LAUNCH_OK();

  ScanTemp<unsigned long long> tmp{v.block_sums, v.counter, v.total};
  VoxHeadFlag flag{v.pslot, v.hfirst, v.hcount};
  VoxAssign assign{v.pslot, v.hkeys, v.hoff, v.vslot, coors, coors_ncol, batch_idx,
                   max_voxels, g.gx, g.gy, v.cutoff};
  MSMD_CUDA_OK((device_exclusive_scan<unsigned long long>(flag, assign, num_points, tmp, stream)));

  vox_bucket_kernel<<<ceil_div(num_points, tpb), tpb, 0, stream>>>(
      num_points, v.pslot, v.hoff, v.hcursor, v.bucket, v.cutoff, v.total, max_voxels,
      voxel_num);
  MSMD_LAUNCH_OK();

  const int cap_voxels = num_points < max_voxels ? num_points : max_voxels;
  const long long threads = (long long)cap_voxels * 32;
  vox_gather_kernel<<<ceil_div(threads, tpb), tpb, 0, stream>>>(
      points, num_features, max_points, max_voxels, v.total, v.vslot, v.hoff, v.hcursor,
      v.bucket, voxels, num_points_per_voxel, mean, mean_features);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}
