// gma.cu -- the gating / concatenation step of the Gated Modality-Aware convolution, one kernel.
//
// Reference (mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py, one stage of grouped_sparse_conv):
//   :371-377  cross_gating = ReLU(Linear(C3 -> 64))(cat[feat3, dummy]);  only2 = cross_gating[nn_idx] * feat2[only-2D]
//   :391-401  mixed = cat[ feat3[syn3],  ReLU(Linear(C3 -> 64))(feat3[syn3]) * feat2[syn2] ]
//   :414-425  unified = cat[ pad(only3_conv_out, (0, 64)),  pad(only2, (C3, 0)),  mixed ]  (+ the coordinates)
// i.e. an nn.Linear over N3+1 rows, two more over the gathered rows, four index_selects, three pads / cats and a
// zero-filled buffer -- ~20 eager kernels and as many Python-level calls per stage in round 1.  Here every row of the
// unified tensor is produced by ONE warp: it gathers the 3-D feature row that gates it (or the dummy embedding for an
// unassigned only-2D voxel, nn_idx = -1), evaluates the 64 gate outputs against the weight matrix held TRANSPOSED in
// shared memory ([C3][64]: lane o and o+32 read consecutive words, no bank conflicts), applies ReLU, multiplies the
// 2-D feature row and writes the row -- zero padding included -- exactly once.  The Linear is evaluated only for rows
// that are used (the reference evaluates it for all N3 + 1 rows and gathers afterwards); the arithmetic per output is
// the same fp32 dot product + bias, in channel order.
#include "common.cuh"

namespace msmd {

constexpr int kGmaThreads = 256;
constexpr int kGmaGate = 64;   // gate width = channels of the 2-D (virtual-point) features

__global__ void __launch_bounds__(kGmaThreads)
gather_rows_kernel(const float* __restrict__ feat, int c, const int4* __restrict__ coords,
                   const long long* __restrict__ rows, int n, float* __restrict__ out_feat,
                   int4* __restrict__ out_coords) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < n; i += nwarps) {
    const long long r = rows[i];
    if (out_feat) {
      const float* src = feat + (size_t)r * c;
      float* dst = out_feat + (size_t)i * c;
      for (int ch = lane; ch < c; ch += 32) dst[ch] = __ldg(src + ch);
    }
    if (lane == 0 && out_coords) out_coords[i] = __ldg(coords + r);
  }
}

__global__ void __launch_bounds__(kGmaThreads)
gma_assemble_kernel(const float* __restrict__ y_only3, const int4* __restrict__ idx_only3, int n_o3,
                    const float* __restrict__ feat3, int c3, const float* __restrict__ feat2,
                    const int4* __restrict__ bz2, const long long* __restrict__ only2_rows,
                    const int4* __restrict__ only2_bzyx, const long long* __restrict__ nn_idx, int n_o2, int rows_o2,
                    const long long* __restrict__ syn3, const long long* __restrict__ syn2, int n_mix, int rows_mix,
                    const float* __restrict__ dummy, const float* __restrict__ w_cross,
                    const float* __restrict__ b_cross, const float* __restrict__ w_gate,
                    const float* __restrict__ b_gate, float* __restrict__ out, int4* __restrict__ out_idx) {
  extern __shared__ uint8_t gma_smem_raw[];
  float* wt_cross = (float*)gma_smem_raw;                       // [c3][64]
  float* wt_gate = wt_cross + c3 * kGmaGate;        // [c3][64]
  float* xrow = wt_gate + c3 * kGmaGate;            // [warps][c3]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  for (int t = threadIdx.x; t < c3 * kGmaGate; t += blockDim.x) {
    const int o = t / c3, ch = t - o * c3;          // nn.Linear weight is [64][c3]
    wt_cross[ch * kGmaGate + o] = __ldg(w_cross + t);
    wt_gate[ch * kGmaGate + o] = __ldg(w_gate + t);
  }
  __syncthreads();
  float* xs = xrow + wib * c3;
  const int cu = c3 + kGmaGate;
  const int total = n_o3 + rows_o2 + rows_mix;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < total; row += nwarps) {
    float* dst = out + (size_t)row * cu;
    if (row < n_o3) {
      // only-3D voxel: [convolved 3-D feature | 0]
      if (out) {
        for (int ch = lane; ch < c3; ch += 32) dst[ch] = __ldg(y_only3 + (size_t)row * c3 + ch);
        dst[c3 + lane] = 0.f;
        dst[c3 + 32 + lane] = 0.f;
      }
      if (lane == 0 && out_idx) out_idx[row] = __ldg(idx_only3 + row);
      continue;
    }
    const bool is_o2 = row < n_o3 + rows_o2;
    const int j = is_o2 ? row - n_o3 : row - n_o3 - rows_o2;
    const bool real = is_o2 ? (j < n_o2 && only2_rows != nullptr) : (j < n_mix);
    // the 3-D row that gates this voxel (mixed: also its own 3-D half) and the 2-D row it scales
    const float* g3 = dummy;
    const float* f2 = nullptr;
    int4 coord = make_int4(0, 0, 0, 0);
    if (is_o2) {
      if (j < n_o2) {
        const long long nn = __ldg(nn_idx + j);
        if (nn >= 0) g3 = feat3 + (size_t)nn * c3;
        coord = __ldg(only2_bzyx + j);
        if (only2_rows) f2 = feat2 + (size_t)__ldg(only2_rows + j) * kGmaGate;
      }
    } else if (real) {
      g3 = feat3 + (size_t)__ldg(syn3 + j) * c3;
      const long long r2 = __ldg(syn2 + j);
      f2 = feat2 + (size_t)r2 * kGmaGate;
      coord = __ldg(bz2 + r2);
    }
    if (lane == 0 && out_idx) out_idx[row] = coord;
    if (!out) continue;   // coordinates only
    if (!real || f2 == nullptr) {   // the all-zero voxel pad_missing_batch_id appends for an empty group (:208-225)
      for (int ch = lane; ch < cu; ch += 32) dst[ch] = 0.f;
      continue;
    }
    __syncwarp();
    for (int ch = lane; ch < c3; ch += 32) xs[ch] = __ldg(g3 + ch);
    __syncwarp();
    const float* wt = is_o2 ? wt_cross : wt_gate;
    const float* bias = is_o2 ? b_cross : b_gate;
    float a0 = __ldg(bias + lane), a1 = __ldg(bias + 32 + lane);
    for (int ch = 0; ch < c3; ++ch) {
      const float x = xs[ch];
      a0 = fmaf(x, wt[ch * kGmaGate + lane], a0);
      a1 = fmaf(x, wt[ch * kGmaGate + 32 + lane], a1);
    }
    a0 = fmaxf(a0, 0.f) * __ldg(f2 + lane);
    a1 = fmaxf(a1, 0.f) * __ldg(f2 + 32 + lane);
    if (is_o2) {
      for (int ch = lane; ch < c3; ch += 32) dst[ch] = 0.f;
    } else {
      for (int ch = lane; ch < c3; ch += 32) dst[ch] = xs[ch];
    }
    dst[c3 + lane] = a0;
    dst[c3 + 32 + lane] = a1;
  }
}

}  // namespace msmd

using namespace msmd;

extern "C" MSMD_API int msmd_gather_rows(const float* features, int channels, const int* coords4,
                                         const long long* rows, int n, float* out_features, int* out_coords4,
                                         msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(n >= 0 && channels > 0, "gather_rows: bad sizes");
  if (n == 0) return MSMD_OK;
  // features / coordinates may be gathered by separate calls (out_features or out_coords4 NULL, not both)
  MSMD_REQUIRE(rows && (out_features || out_coords4) && (!out_features || features) && (!out_coords4 || coords4),
               "gather_rows: null pointer");
  int blocks = ceil_div((long long)n * 32, kGmaThreads);
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  gather_rows_kernel<<<blocks, kGmaThreads, 0, stream>>>(features, channels, (const int4*)coords4, rows, n,
                                                        out_features, (int4*)out_coords4);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_gma_assemble(const float* y_only3, const int* idx_only3, int n_only3,
                                          const float* feat3, int n3, int c3, const float* feat2, const int* bz2,
                                          int n2, int c2, const long long* only2_rows, const int* only2_bzyx,
                                          const long long* nn_idx, int n_only2, const long long* syn3,
                                          const long long* syn2, int n_mix, const float* dummy, const float* w_cross,
                                          const float* b_cross, const float* w_gate, const float* b_gate,
                                          float* unified_features, int* unified_indices, msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(c2 == kGmaGate, "gma_assemble: the 2-D features must have 64 channels (got %d)", c2);
  MSMD_REQUIRE(c3 >= 1 && c3 <= 256 && n_only3 >= 0 && n_only2 >= 0 && n_mix >= 0 && n3 >= 0 && n2 >= 0,
               "gma_assemble: bad sizes");
  MSMD_REQUIRE((unified_features || unified_indices) && dummy && w_cross && b_cross && w_gate && b_gate,
               "gma_assemble: null pointer");
  MSMD_REQUIRE(n_only3 == 0 || ((y_only3 || !unified_features) && (idx_only3 || !unified_indices)),
               "gma_assemble: only-3D rows without data");
  MSMD_REQUIRE(n_only2 == 0 || (only2_bzyx && nn_idx), "gma_assemble: only-2D rows without data");
  MSMD_REQUIRE(n_mix == 0 || (syn3 && syn2 && feat3 && feat2 && bz2), "gma_assemble: mixed rows without data");
  // an empty only-2D / mixed group still contributes one all-zero voxel (pad_missing_batch_id, one sample per GPU)
  const int rows_o2 = n_only2 > 0 ? n_only2 : 1, rows_mix = n_mix > 0 ? n_mix : 1;
  const long long total = (long long)n_only3 + rows_o2 + rows_mix;
  const size_t smem = ((size_t)2 * c3 * kGmaGate + (size_t)(kGmaThreads / 32) * c3) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    MSMD_CUDA_OK(cudaFuncSetAttribute(gma_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  int blocks = ceil_div(total * 32, kGmaThreads);
  if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;   // each CTA stages both weight matrices once
  gma_assemble_kernel<<<blocks, kGmaThreads, smem, stream>>>(
      y_only3, (const int4*)idx_only3, n_only3, feat3, c3, feat2, (const int4*)bz2, only2_rows,
      (const int4*)only2_bzyx, nn_idx, n_only2, rows_o2, syn3, syn2, n_mix, rows_mix, dummy, w_cross, b_cross, w_gate,
      b_gate, unified_features, (int4*)unified_indices);
  MSMD_LAUNCH_OK();
  return MSMD_OK;
}

namespace {
struct StageArena {
  char* base;
  size_t size, used;
  template <typename T>
  T* take(size_t count) {
    const size_t off = msmd::align_up(used, 256);
    const size_t end = off + sizeof(T) * (count ? count : 1);
    used = end;
    if (end > size) return nullptr;
    return (T*)(base + off);
  }
};
}  // namespace

#define MSMD_STAGE_TAKE(ptr, T, count)                                                          \
  T* ptr = arena.take<T>(count);                                                                \
  if (!ptr) {                                                                                   \
    set_error("gma_stage_forward: arena too small (%zu bytes needed so far, %zu given)",        \
              arena.used, arena.size);                                                          \
    return MSMD_ERR_WORKSPACE;                                                                  \
  }
#define MSMD_STAGE_TRY(expr)          \
  do {                                \
    int _r = (expr);                  \
    if (_r != MSMD_OK) return _r;     \
  } while (0)

// a conv chain on the remaining part of the stage's arena; `last` = the chain's final activation.  The chain's input
// indices were produced on the geometry stream (MSMD_NET_INDICES_ON_GEOMETRY_STREAM).
static int run_chain(const msmd_conv_layer* layers, int n_layers, const float* feat, const int* idx, int n, int c,
                     int batch, const int* shape, StageArena& arena, msmd_sparse_desc* last, cudaStream_t stream) {
  msmd_sparse_desc acts[8];
  MSMD_REQUIRE(n_layers >= 1 && n_layers < 8, "gma_stage_forward: a chain has 1..7 layers");
  const size_t off = align_up(arena.used, 256);
  if (off >= arena.size) {
    set_error("gma_stage_forward: arena too small");
    return MSMD_ERR_WORKSPACE;
  }
  size_t used = 0;
  MSMD_STAGE_TRY(msmd_sparse_net_forward_ex(layers, n_layers, feat, idx, n, c, batch, shape, arena.base + off,
                                            arena.size - off, acts, &used, MSMD_NET_INDICES_ON_GEOMETRY_STREAM,
                                            (msmd_stream_t)stream));
  arena.used = off + used;
  *last = acts[n_layers];
  return MSMD_OK;
}

static cudaEvent_t g_stage_event[16] = {};   // per device: "the union grid of sparse_add is on the geometry stream"

extern "C" MSMD_API int msmd_gma_stage_forward(const msmd_gma_stage* st, const float* feat3, const int* bz3, int n3,
                                               const float* feat2, const int* bz2, int n2,
                                               const long long* only3_rows, int n_only3,
                                               const long long* only2_rows, const int* only2_bzyx,
                                               const long long* nn_idx, int n_only2, const long long* syn3,
                                               const long long* syn2, int n_mix, const float* dummy,
                                               const float* prev_features, const int* prev_indices, int n_prev,
                                               int batch_size, const int* shape, void* arena_ptr, size_t arena_bytes,
                                               msmd_sparse_desc* out, void* indices_ready_event,
                                               msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MSMD_REQUIRE(st && out && arena_ptr && shape && batch_size == 1, "gma_stage_forward: bad arguments (one sample per GPU)");
  MSMD_REQUIRE(n3 > 0 && n_only3 > 0 && feat3 && bz3 && only3_rows, "gma_stage_forward: needs 3-D voxels");
  void* geom_v = nullptr;
  MSMD_STAGE_TRY(msmd_executor_geometry_stream(&geom_v));
  cudaStream_t geom = (cudaStream_t)geom_v;
  int dev = 0;
  MSMD_CUDA_OK(cudaGetDevice(&dev));
  MSMD_REQUIRE(dev >= 0 && dev < 16, "gma_stage_forward: device ordinal %d unsupported", dev);
  if (!g_stage_event[dev]) MSMD_CUDA_OK(cudaEventCreateWithFlags(&g_stage_event[dev], cudaEventDisableTiming));
  // what the geometry stream reads (row lists, 2-D coordinates) is complete behind `indices_ready_event`; the LiDAR
  // index sets and the previous stage's coordinates were produced on the geometry stream itself
  if (indices_ready_event) {
    MSMD_CUDA_OK(cudaStreamWaitEvent(geom, (cudaEvent_t)indices_ready_event, 0));
  } else {
    MSMD_CUDA_OK(cudaEventRecord(g_stage_event[dev], stream));
    MSMD_CUDA_OK(cudaStreamWaitEvent(geom, g_stage_event[dev], 0));
  }
  StageArena arena{(char*)arena_ptr, arena_bytes, 0};
  const int c3 = st->c3, cu = st->c3 + st->c2;
  const msmd_stream_t geom_ = (msmd_stream_t)geom;
  // 1. only-3D voxels -> their SubM chain (coordinates on the geometry stream, features on the caller's)
  MSMD_STAGE_TAKE(f_o3, float, (size_t)n_only3 * c3);
  MSMD_STAGE_TAKE(i_o3, int, (size_t)n_only3 * 4);
  MSMD_STAGE_TRY(msmd_gather_rows(nullptr, c3, bz3, only3_rows, n_only3, nullptr, i_o3, geom_));
  MSMD_STAGE_TRY(msmd_gather_rows(feat3, c3, nullptr, only3_rows, n_only3, f_o3, nullptr, stream_));
  msmd_sparse_desc y3;
  MSMD_STAGE_TRY(run_chain(st->only3d, st->n_only3d, f_o3, i_o3, n_only3, c3, batch_size, shape, arena, &y3, stream));
  MSMD_REQUIRE(y3.n == n_only3 && y3.channels == c3, "gma_stage_forward: the only-3D chain must keep rows and channels");
  // 2. the unified voxel list: coordinates (geometry stream), gates + zero-padded concatenation (caller's stream)
  const int n_uni = n_only3 + (n_only2 > 0 ? n_only2 : 1) + (n_mix > 0 ? n_mix : 1);
  MSMD_STAGE_TAKE(uf, float, (size_t)n_uni * cu);
  MSMD_STAGE_TAKE(ui, int, (size_t)n_uni * 4);
  MSMD_STAGE_TRY(msmd_gma_assemble(nullptr, i_o3, n_only3, feat3, n3, c3, feat2, bz2, n2, st->c2, only2_rows,
                                   only2_bzyx, nn_idx, n_only2, syn3, syn2, n_mix, dummy, st->w_cross, st->b_cross,
                                   st->w_gate, st->b_gate, nullptr, ui, geom_));
  MSMD_STAGE_TRY(msmd_gma_assemble(y3.features, nullptr, n_only3, feat3, n3, c3, feat2, bz2, n2, st->c2, only2_rows,
                                   only2_bzyx, nn_idx, n_only2, syn3, syn2, n_mix, dummy, st->w_cross, st->b_cross,
                                   st->w_gate, st->b_gate, uf, nullptr, stream_));
  // 3. aggregation block
  msmd_sparse_desc agg;
  MSMD_STAGE_TRY(run_chain(st->agg, st->n_agg, uf, ui, n_uni, cu, batch_size, shape, arena, &agg, stream));
  // 4. + the previous stage's output (Fsp.sparse_add, :455): union grid, count and coordinates on the geometry
  //    stream (the SubM aggregation block keeps the unified coordinates), the feature sum on the caller's
  const float* sf = agg.features;
  const int* si = ui;
  int sn = n_uni;
  if (prev_features) {
    MSMD_REQUIRE(prev_indices && n_prev >= 0, "gma_stage_forward: previous stage without indices");
    const size_t words = msmd_grid_num_words(batch_size, shape);
    MSMD_STAGE_TAKE(bits, uint32_t, words);
    MSMD_STAGE_TAKE(prefix, int, words);
    MSMD_STAGE_TAKE(count, int, 1);
    const size_t ws_bytes = msmd_scan_workspace();
    MSMD_STAGE_TAKE(ws, char, ws_bytes);
    MSMD_STAGE_TRY(msmd_sparse_add_outputs(ui, n_uni, prev_indices, n_prev, batch_size, shape, bits, prefix, count, ws,
                                           ws_bytes, geom_));
    int h_count = 0;
    MSMD_CUDA_OK(cudaMemcpyAsync(&h_count, count, sizeof(int), cudaMemcpyDeviceToHost, geom));
    MSMD_CUDA_OK(cudaStreamSynchronize(geom));   // drains the geometry stream only
    MSMD_STAGE_TAKE(of, float, (size_t)h_count * cu);
    MSMD_STAGE_TAKE(oi, int, (size_t)h_count * 4);
    MSMD_STAGE_TRY(msmd_sparse_add_finish(bits, prefix, h_count, ui, nullptr, n_uni, prev_indices, nullptr, n_prev, cu,
                                          batch_size, shape, oi, nullptr, geom_));
    MSMD_CUDA_OK(cudaEventRecord(g_stage_event[dev], geom));
    MSMD_CUDA_OK(cudaStreamWaitEvent(stream, g_stage_event[dev], 0));
    MSMD_STAGE_TRY(msmd_sparse_add_finish(bits, prefix, h_count, ui, agg.features, n_uni, prev_indices, prev_features,
                                          n_prev, cu, batch_size, shape, nullptr, of, stream_));
    sf = of; si = oi; sn = h_count;
  }
  // 5. downscale convolution
  MSMD_STAGE_TRY(run_chain(st->down, st->n_down, sf, si, sn, cu, batch_size, shape, arena, out, stream));
  return MSMD_OK;
}
