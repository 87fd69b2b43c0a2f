// executor.cu -- native executor for a chain of sparse convolutions (the host-side loop of
// SparseEncoder.forward, mmdet3d/models/middle_encoders/sparse_encoder.py:96-133, and of the
// spconv SparseSequential / SparseBasicBlock modules it is made of).
//
// The Python modules of this package launch one C-ABI call per rulebook / convolution, i.e.
// ~55 ctypes calls + allocations + 4 blocking size read-backs per LiDAR scene: ~1.2 ms of host
// time against ~1.9 ms of GPU time.  This executor runs the same sequence from C++ with ONE call:
// the layer list ("plan") is built once from the module tree, all intermediates are bump-allocated
// from a caller-provided arena (the library still never allocates), and only the N_out of the
// strided convolutions is read back (the output tensors cannot be sized without it).
// Results are bit-identical to the module path -- the same kernels run on the same operands.
//
// Two streams: rulebooks depend on coordinates only, so the geometry chain (bit grids, rulebooks,
// the N_out read-backs) runs on a library-owned auxiliary stream while the feature convolutions
// run on the caller's stream, joined by events.  The host blocks only on the geometry stream --
// a few tiny kernels -- while the convolutions of the previous resolution level keep the GPU busy,
// so the read-backs no longer drain the pipeline.  (One process per GPU, single host thread.)
#include <map>
#include <vector>

#include "common.cuh"

namespace msmd {

struct RulebookX {
  int* pair = nullptr;
  int* pair_sorted = nullptr;   // mask-sorted copy of the table + its slot -> row map (opt-in)
  int* row_perm = nullptr;
  unsigned* tile_mask = nullptr;   // msmd_rulebook_tile_masks of the table the split-operand kernel tiles
  cudaEvent_t ready = nullptr;  // recorded on the geometry stream after the rulebook kernels
  int seq = -1;                 // position of `ready` among the events recorded on the geometry stream
};

static int g_mask_sort = 0;  // msmd_spconv_set_mask_sort

struct IndexSetX {
  int* indices = nullptr;
  int n = 0;
  int shape[3] = {0, 0, 0};
  uint32_t* bits = nullptr;
  int* prefix = nullptr;
  int* perm = nullptr;
  bool has_grid = false;
  bool ordered = false;  // rows already in ascending linear order (output of a strided conv)
  std::map<std::vector<int>, RulebookX> subm;  // (ksize, dilation) -> pair_fwd
};

struct Arena {
  char* base;
  size_t size, used;
  template <typename T>
  T* take(size_t count) {
    const size_t off = align_up(used, 256);
    const size_t end = off + sizeof(T) * (count ? count : 1);
    used = end;
    if (end > size) return nullptr;
    return (T*)(base + off);
  }
};

// library-owned auxiliary stream + event pool (per device, created on first use)
struct AuxStreams {
  cudaStream_t geom = nullptr;
  std::vector<cudaEvent_t> events;
  size_t next = 0;
};
static AuxStreams g_aux[16];

// The geometry stream carries the short kernels everything else waits for (index sets, rulebooks, tile masks) while
// the convolutions keep every SM busy: it gets the highest stream priority, so that its thread blocks are the first to
// be placed whenever an SM frees up (MSMD_GEOM_PRIORITY=0 in the environment: default priority, for A/B runs).
static int create_geometry_stream(cudaStream_t* out) {
  int least = 0, greatest = 0;
  MSMD_CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  const char* e = getenv("MSMD_GEOM_PRIORITY");
  const int prio = (e && e[0] == '0') ? least : greatest;
  MSMD_CUDA_OK(cudaStreamCreateWithPriority(out, cudaStreamNonBlocking, prio));
  return MSMD_OK;
}

static int aux_for_current_device(AuxStreams** out) {
  int dev = 0;
  MSMD_CUDA_OK(cudaGetDevice(&dev));
  MSMD_REQUIRE(dev >= 0 && dev < 16, "sparse_net_forward: device ordinal %d unsupported", dev);
  AuxStreams& a = g_aux[dev];
  if (!a.geom) {
    const int rc = create_geometry_stream(&a.geom);
    if (rc != MSMD_OK) return rc;
  }
  a.next = 0;
  *out = &a;
  return MSMD_OK;
}

static int next_event(AuxStreams& a, cudaEvent_t* ev) {
  if (a.next == a.events.size()) {
    cudaEvent_t e;
    MSMD_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    a.events.push_back(e);
  }
  *ev = a.events[a.next++];
  return MSMD_OK;
}

}  // namespace msmd

using namespace msmd;

#define MSMD_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != MSMD_OK) return _r; \
  } while (0)

#define MSMD_ARENA(ptr, T, count)                                                              \
  T* ptr = arena.take<T>(count);                                                               \
  if (!ptr) {                                                                                  \
    set_error("sparse_net_forward: arena too small (%zu bytes needed so far, %zu given)",      \
              arena.used, arena.size);                                                         \
    return MSMD_ERR_WORKSPACE;                                                                 \
  }

static int ensure_grid(IndexSetX& s, int batch, Arena& arena, void* scan_ws, size_t scan_ws_bytes,
                       cudaStream_t stream) {
  if (s.has_grid) return MSMD_OK;
  const size_t words = msmd_grid_num_words(batch, s.shape);
  MSMD_ARENA(bits, uint32_t, words);
  MSMD_ARENA(prefix, int, words);
  int* perm = nullptr;
  if (!s.ordered) {
    MSMD_ARENA(p, int, (size_t)s.n);
    perm = p;
  }
  MSMD_ARENA(count, int, 1);
  MSMD_TRY(msmd_grid_build(s.indices, s.n, batch, s.shape, bits, prefix, perm, count, scan_ws,
                           scan_ws_bytes, (msmd_stream_t)stream));
  s.bits = bits; s.prefix = prefix; s.perm = perm; s.has_grid = true;
  return MSMD_OK;
}

// The library-owned geometry stream of the current device (created on first use).  A caller that wants to start
// work which depends on an executor call's COORDINATES only (index sets, rulebooks) -- not on its features -- records
// an event on this stream right after the call returns and waits for that instead of for its own stream.
extern "C" MSMD_API int msmd_executor_geometry_stream(void** stream_out) {
  MSMD_REQUIRE(stream_out, "executor_geometry_stream: null argument");
  int dev = 0;
  MSMD_CUDA_OK(cudaGetDevice(&dev));
  MSMD_REQUIRE(dev >= 0 && dev < 16, "executor_geometry_stream: device ordinal %d unsupported", dev);
  AuxStreams* aux = &g_aux[dev];
  if (!aux->geom) {
    const int rc = create_geometry_stream(&aux->geom);
    if (rc != MSMD_OK) return rc;
  }
  *stream_out = (void*)aux->geom;
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_spconv_set_mask_sort(int enable) {
  g_mask_sort = enable ? 1 : 0;
  return MSMD_OK;
}

extern "C" MSMD_API int msmd_sparse_net_forward(const msmd_conv_layer* layers, int n_layers,
                                                const float* features, const int* indices, int n,
                                                int channels, int batch_size, const int* spatial_shape,
                                                void* arena_ptr, size_t arena_bytes,
                                                msmd_sparse_desc* acts, msmd_stream_t stream_) {
  return msmd_sparse_net_forward_ex(layers, n_layers, features, indices, n, channels, batch_size, spatial_shape,
                                    arena_ptr, arena_bytes, acts, nullptr, 0, stream_);
}

extern "C" MSMD_API int msmd_sparse_net_forward_ex(const msmd_conv_layer* layers, int n_layers,
                                                   const float* features, const int* indices, int n,
                                                   int channels, int batch_size, const int* spatial_shape,
                                                   void* arena_ptr, size_t arena_bytes,
                                                   msmd_sparse_desc* acts, size_t* arena_used, int flags,
                                                   msmd_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (arena_used) *arena_used = 0;
  MSMD_REQUIRE(layers && n_layers > 0 && acts && arena_ptr, "sparse_net_forward: null argument");
  MSMD_REQUIRE(n >= 0 && channels > 0 && batch_size > 0, "sparse_net_forward: bad input sizes");
  Arena arena{(char*)arena_ptr, arena_bytes, 0};
  const size_t scan_ws_bytes = msmd_scan_workspace();
  MSMD_ARENA(scan_ws, char, scan_ws_bytes);

  AuxStreams* aux = nullptr;
  MSMD_TRY(aux_for_current_device(&aux));
  cudaStream_t geom = aux->geom;
  if (!(flags & MSMD_NET_INDICES_ON_GEOMETRY_STREAM)) {
    // the geometry stream starts after everything already queued on the caller's stream (which produced `indices`);
    // with the flag the caller produced them ON the geometry stream, and the rulebooks may run ahead of the caller's
    // stream -- i.e. ahead of the previous chain's convolutions -- as they do inside one call
    cudaEvent_t ev_in;
    MSMD_TRY(next_event(*aux, &ev_in));
    MSMD_CUDA_OK(cudaEventRecord(ev_in, stream));
    MSMD_CUDA_OK(cudaStreamWaitEvent(geom, ev_in, 0));
  }

  std::vector<IndexSetX> isets;
  isets.reserve(n_layers + 1);
  std::vector<int> act_iset(n_layers + 1, -1);
  // split-bf16 operand cache (weight_tc 4, csrc/spconv_sb.cu): split image of activation a, or null.  It is written
  // by the producing layer's epilogue when some later layer of the plan gathers the activation through that path.
  std::vector<void*> act_split(n_layers + 1, nullptr);
  std::vector<char> wants_split(n_layers + 1, 0);
  for (int li = 0; li < n_layers; ++li)
    if (layers[li].weight_tc == 4 && layers[li].input >= 0 && layers[li].input <= li) wants_split[layers[li].input] = 1;
  {
    IndexSetX s0;
    s0.indices = (int*)indices; s0.n = n;
    for (int d = 0; d < 3; ++d) s0.shape[d] = spatial_shape[d];
    isets.push_back(s0);
    act_iset[0] = 0;
    acts[0].features = (float*)features; acts[0].indices = (int*)indices; acts[0].n = n;
    acts[0].channels = channels;
    for (int d = 0; d < 3; ++d) acts[0].spatial_shape[d] = spatial_shape[d];
  }

  int rc = MSMD_OK;
  int ready_seq = 0, waited_seq = 0;  // events recorded on the geometry stream / the latest one `stream` waited for
  for (int li = 0; li < n_layers && rc == MSMD_OK; ++li) {
    rc = [&]() -> int {
      const msmd_conv_layer& L = layers[li];
      MSMD_REQUIRE(L.input >= 0 && L.input <= li, "sparse_net_forward: layer %d reads activation %d", li,
                   L.input);
      MSMD_REQUIRE(L.residual < 0 || L.residual <= li, "sparse_net_forward: layer %d bad residual", li);
      const msmd_sparse_desc in = acts[L.input];
      MSMD_REQUIRE(in.channels == L.cin, "sparse_net_forward: layer %d expects %d channels, got %d", li,
                   L.cin, in.channels);
      const int kvol = L.ksize[0] * L.ksize[1] * L.ksize[2];
      const int in_id = act_iset[L.input];
      MSMD_TRY(ensure_grid(isets[in_id], batch_size, arena, scan_ws, scan_ws_bytes, geom));

      int out_id, n_out;
      RulebookX rb;
      if (L.subm) {
        IndexSetX& s = isets[in_id];
        std::vector<int> key(L.ksize, L.ksize + 3);
        key.insert(key.end(), L.dilation, L.dilation + 3);
        auto it = s.subm.find(key);
        if (it == s.subm.end()) {
          MSMD_ARENA(p, int, (size_t)kvol * (size_t)s.n);
          MSMD_TRY(msmd_rulebook_subm(s.indices, s.n, batch_size, s.shape, L.ksize, L.dilation, s.bits,
                                      s.prefix, s.perm, p, (msmd_stream_t)geom));
          rb.pair = p;
          if (g_mask_sort && kvol == 27 && s.n > 0) {
            // group the rows by neighbour-mask structure once; every SubM layer on this index set
            // (4-5 of them) then tiles the permuted table
            const size_t sort_bytes = msmd_rulebook_mask_sort_workspace(s.n);
            MSMD_ARENA(rp, int, (size_t)s.n);
            MSMD_ARENA(ps, int, (size_t)kvol * (size_t)s.n);
            MSMD_ARENA(sort_ws, char, sort_bytes);
            MSMD_TRY(msmd_rulebook_mask_sort(p, kvol, s.n, rp, ps, sort_ws, sort_bytes, (msmd_stream_t)geom));
            rb.row_perm = rp;
            rb.pair_sorted = ps;
          }
          if (L.weight_tc == 4 && msmd_spconv_sb_uses_tile_masks() && s.n > 0) {
            MSMD_ARENA(tm, unsigned, (size_t)((s.n + 127) / 128));
            MSMD_TRY(msmd_rulebook_tile_masks(rb.pair_sorted ? rb.pair_sorted : rb.pair, kvol, s.n, tm,
                                              (msmd_stream_t)geom));
            rb.tile_mask = tm;
          }
          MSMD_TRY(next_event(*aux, &rb.ready));
          MSMD_CUDA_OK(cudaEventRecord(rb.ready, geom));
          rb.seq = ++ready_seq;
          s.subm[key] = rb;
        } else {
          rb = it->second;
        }
        out_id = in_id;
        n_out = s.n;
      } else {
        const IndexSetX s = isets[in_id];  // copy: isets may reallocate below
        IndexSetX o;
        MSMD_TRY(msmd_conv_out_shape(s.shape, L.ksize, L.stride, L.padding, L.dilation, o.shape));
        const size_t words = msmd_grid_num_words(batch_size, o.shape);
        MSMD_ARENA(obits, uint32_t, words);
        MSMD_ARENA(oprefix, int, words);
        MSMD_ARENA(count, int, 1);
        MSMD_TRY(msmd_rulebook_conv_outputs(s.indices, s.n, batch_size, s.shape, L.ksize, L.stride,
                                            L.padding, L.dilation, obits, oprefix, count, scan_ws,
                                            scan_ws_bytes, (msmd_stream_t)geom));
        // the one unavoidable read-back: N_out sizes the output rows and the pair table.  Only the
        // geometry stream is drained; the convolutions queued on the caller's stream keep running.
        int h_count = 0;
        MSMD_CUDA_OK(cudaMemcpyAsync(&h_count, count, sizeof(int), cudaMemcpyDeviceToHost, geom));
        MSMD_CUDA_OK(cudaStreamSynchronize(geom));
        n_out = h_count;
        MSMD_ARENA(oidx, int, (size_t)4 * (size_t)n_out);
        MSMD_ARENA(p, int, (size_t)kvol * (size_t)n_out);
        MSMD_TRY(msmd_rulebook_conv_pairs(obits, oprefix, n_out, batch_size, s.shape, L.ksize, L.stride,
                                          L.padding, L.dilation, s.bits, s.prefix, s.perm, oidx, p,
                                          (msmd_stream_t)geom));
        rb.pair = p;
        if (L.weight_tc == 4 && msmd_spconv_sb_uses_tile_masks() && n_out > 0) {
          MSMD_ARENA(tm, unsigned, (size_t)((n_out + 127) / 128));
          MSMD_TRY(msmd_rulebook_tile_masks(p, kvol, n_out, tm, (msmd_stream_t)geom));
          rb.tile_mask = tm;
        }
        MSMD_TRY(next_event(*aux, &rb.ready));
        MSMD_CUDA_OK(cudaEventRecord(rb.ready, geom));
        rb.seq = ++ready_seq;
        o.indices = oidx; o.n = n_out; o.bits = obits; o.prefix = oprefix; o.perm = nullptr;
        o.has_grid = true; o.ordered = true;
        isets.push_back(o);
        out_id = (int)isets.size() - 1;
      }

      MSMD_ARENA(out, float, (size_t)n_out * (size_t)L.cout);
      const float* residual = nullptr;
      if (L.residual >= 0) {
        MSMD_REQUIRE(acts[L.residual].n == n_out && acts[L.residual].channels == L.cout,
                     "sparse_net_forward: layer %d residual shape mismatch", li);
        residual = acts[L.residual].features;
      }
      // rulebook (and its indices) ready.  The geometry stream is in-order, so having waited for a later
      // event covers every earlier one: the 4-5 layers that share a rulebook wait once, not once each
      if (rb.seq > waited_seq) {
        MSMD_CUDA_OK(cudaStreamWaitEvent(stream, rb.ready, 0));
        waited_seq = rb.seq;
      }
      if (L.weight_tc == 4) {  // bf16x3 through the split-bf16 operand cache
        if (!act_split[L.input] && in.n > 0) {   // network input, or produced by a layer of another kind
          MSMD_ARENA(xs, uint16_t, (size_t)in.n * (size_t)msmd_split_width(L.cin));
          MSMD_TRY(msmd_split_bf16(in.features, in.n, L.cin, xs, (msmd_stream_t)stream));
          act_split[L.input] = xs;
        }
        void* out_s = nullptr;
        if (wants_split[li + 1] && n_out > 0) {
          MSMD_ARENA(os, uint16_t, (size_t)n_out * (size_t)msmd_split_width(L.cout));
          out_s = os;
        }
        MSMD_TRY(msmd_spconv_fwd_sb_ex(act_split[L.input], in.n, L.weight, rb.pair_sorted ? rb.pair_sorted : rb.pair,
                                       rb.pair_sorted ? rb.row_perm : nullptr, rb.tile_mask, n_out, L.cin, L.cout, kvol,
                                       L.scale, L.shift, residual, L.relu, out, out_s, (msmd_stream_t)stream));
        act_split[li + 1] = out_s;
      } else if (L.weight_tc == 2 || L.weight_tc == 3) {  // 16-bit operand kernels: bf16x3 / bf16
        const size_t ws_bytes = msmd_spconv_tc16_workspace(n_out, L.cout);  // variant 3: split-K hand-off buffer
        char* ws = nullptr;
        if (ws_bytes) {
          MSMD_ARENA(w, char, ws_bytes);
          ws = w;
        }
        MSMD_TRY(msmd_spconv_fwd_tc16_ws(in.features, in.n, L.weight, rb.pair_sorted ? rb.pair_sorted : rb.pair,
                                         rb.pair_sorted ? rb.row_perm : nullptr, n_out, L.cin, L.cout, kvol,
                                         L.weight_tc == 2, L.scale, L.shift, residual, L.relu, out, ws, ws_bytes,
                                         (msmd_stream_t)stream));
      } else if (L.weight_tc) {
        const size_t ws_bytes = msmd_spconv_tc_workspace(n_out, L.cout);  // split-K hand-off buffer
        char* ws = nullptr;
        if (ws_bytes) {
          MSMD_ARENA(w, char, ws_bytes);
          ws = w;
        }
        if (rb.pair_sorted)
          MSMD_TRY(msmd_spconv_fwd_tc_sorted(in.features, in.n, L.weight, rb.pair_sorted, rb.row_perm, n_out,
                                             L.cin, L.cout, kvol, L.scale, L.shift, residual, L.relu, out, ws,
                                             ws_bytes, (msmd_stream_t)stream));
        else
          MSMD_TRY(msmd_spconv_fwd_tc_ws(in.features, in.n, L.weight, rb.pair, n_out, L.cin, L.cout, kvol,
                                         L.scale, L.shift, residual, L.relu, out, ws, ws_bytes,
                                         (msmd_stream_t)stream));
      }
      else
        MSMD_TRY(msmd_spconv_fwd(in.features, in.n, L.weight, rb.pair, n_out, L.cin, L.cout, kvol, L.scale,
                                 L.shift, residual, L.relu, out, (msmd_stream_t)stream));
      msmd_sparse_desc& A = acts[li + 1];
      A.features = out; A.indices = isets[out_id].indices; A.n = n_out; A.channels = L.cout;
      for (int d = 0; d < 3; ++d) A.spatial_shape[d] = isets[out_id].shape[d];
      act_iset[li + 1] = out_id;
      return MSMD_OK;
    }();
  }
  // join: whatever the geometry stream still has queued (or an early error exit left behind) is
  // ordered before anything the caller enqueues next on its stream
  cudaEvent_t ev_out;
  if (next_event(*aux, &ev_out) == MSMD_OK) {
    cudaEventRecord(ev_out, geom);
    cudaStreamWaitEvent(stream, ev_out, 0);
  }
  if (arena_used) *arena_used = arena.used;
  return rc;
}
