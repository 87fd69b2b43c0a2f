/*
 * msmd_b200.h -- C ABI of libmsmd_b200.so: the B200 (sm_100a) implementation of the
 * MSMDFusion voxel-space fusion hot path.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless a comment
 *     says "host".  No torch types cross this boundary.
 *   - operands, outputs and scratch ("workspace") are provided by the caller; every op with
 *     scratch has a *_workspace() size query.  The library owns three things, created on first
 *     use: the executor's geometry stream + event pool (per device) and the partial-sum hand-off
 *     slots of the persistent convolution (<= 19 MB per (device, stream); see
 *     msmd_spconv_sb_set_variant).
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); the single-kernel
 *     entries never synchronise.  Counts produced on the device (voxel_num, N_out) are written
 *     to a device int the caller reads back when it needs the value on the host; the two
 *     multi-layer entries (msmd_sparse_net_forward, msmd_gma_stage_forward) read N_out of their
 *     strided convolutions themselves and drain ONLY their geometry stream for it.
 *   - return value: 0 (MSMD_OK) or a negative msmd_status; msmd_last_error() returns a
 *     thread-local message for the last failure.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *     with MSMD_ERR_CUDA.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef MSMD_B200_H_
#define MSMD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSMD_ABI_VERSION 1

#if defined(__GNUC__)
#define MSMD_API __attribute__((visibility("default")))
#else
#define MSMD_API
#endif

typedef enum msmd_status {
  MSMD_OK = 0,
  MSMD_ERR_INVALID = -1,   /* bad argument */
  MSMD_ERR_CUDA = -2,      /* CUDA runtime error (message has the details) */
  MSMD_ERR_WORKSPACE = -3  /* workspace too small / missing */
} msmd_status;

typedef void* msmd_stream_t; /* cudaStream_t */

MSMD_API const char* msmd_last_error(void);
MSMD_API int msmd_abi_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
MSMD_API unsigned long long msmd_launch_count(void);

/* ------------------------------------------------------------------------------------
 * hard_voxelize (+ fused HardSimpleVFE mean)
 *   replaces  voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel,
 *             voxel_size, coors_range, max_points, max_voxels, NDim=3) -> int
 *             mmdet3d/ops/voxel/src/voxelization.h:61-78, CPU kernel
 *             src/voxelization_cpu.cpp:43-142, CUDA src/voxelization_cuda.cu:184-326,
 *   and, when `mean` is given, HardSimpleVFE.forward
 *             mmdet3d/models/voxel_encoders/voxel_encoder.py:30-47.
 *
 *   points  (num_points, num_features) f32
 *   voxels  (cap, max_points, num_features) f32 or NULL.  Rows [0, voxel_num) are fully
 *           written (unused slots zeroed) -- no pre-zeroing needed; rows beyond are untouched.
 *   coors   (cap, coors_ncol) i32, coors_ncol 3 -> (z,y,x); 4 -> (batch_idx,z,y,x)
 *   num_points_per_voxel (cap) i32
 *   mean    (cap, mean_features) f32 or NULL: sum of the first mean_features columns over
 *           the voxel's points / count
 *   voxel_num  device int, receives the number of voxels
 *   cap = min(num_points, max_voxels) rows must be allocated.
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_hard_voxelize_workspace(int num_points);
MSMD_API int msmd_hard_voxelize(const float* points, int num_points, int num_features,
                       const float* voxel_size /* host [3] x,y,z */,
                       const float* coors_range /* host [6] */, int max_points,
                       int max_voxels, float* voxels, int* coors, int coors_ncol,
                       int batch_idx, int* num_points_per_voxel, float* mean,
                       int mean_features, int* voxel_num, void* workspace,
                       size_t workspace_bytes, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Occupancy bit grid: the voxel index structure behind the rulebooks.
 *   One bit per cell of the (batch, D, H, W) grid in linear order ((b*D+z)*H+y)*W+x,
 *   `prefix[w]` = number of set bits in words [0, w), `perm[rank]` = row of the active
 *   voxel with that ascending-linear-order rank (NULL when rows are already in that order).
 *   Plays the role of spconv-2.x's hash table inside ops.get_indice_pairs_implicit_gemm
 *   (call site bug_fix/conv.py:382-396).
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_grid_num_words(int batch_size, const int* spatial_shape /* host [3] */);
MSMD_API size_t msmd_scan_workspace(void);
MSMD_API int msmd_grid_build(const int* indices /* (n,4) b,z,y,x */, int n, int batch_size,
                    const int* spatial_shape /* host [3] */, uint32_t* bits, int* prefix,
                    int* perm /* (n) or NULL */, int* num_active /* device int */,
                    void* workspace, size_t workspace_bytes, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Rulebooks -- replace ops.get_indice_pairs_implicit_gemm (bug_fix/conv.py:382-415).
 *   pair_fwd is (kvol, n_out) i32, kernel offset k = (kz*KY+ky)*KX+kx, -1 = no input.
 * SubM: output rows == input rows.
 * Regular conv: outputs ascending by linear index over the OUTPUT shape.
 *   step 1 (outputs): marks candidate outputs in out_bits, scans -> out_prefix, *num_out
 *   step 2 (pairs):   needs num_out on the host (to size out_indices / pair_fwd)
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_rulebook_subm(const int* indices, int n, int batch_size, const int* spatial_shape,
                       const int* ksize, const int* dilation /* host [3] each */,
                       const uint32_t* bits, const int* prefix, const int* perm,
                       int* pair_fwd, msmd_stream_t stream);
MSMD_API int msmd_conv_out_shape(const int* spatial_shape, const int* ksize, const int* stride,
                        const int* padding, const int* dilation, int* out_shape /* host */);
MSMD_API int msmd_rulebook_conv_outputs(const int* indices, int n, int batch_size,
                               const int* spatial_shape, const int* ksize, const int* stride,
                               const int* padding, const int* dilation, uint32_t* out_bits,
                               int* out_prefix, int* num_out /* device int */,
                               void* workspace, size_t workspace_bytes, msmd_stream_t stream);
MSMD_API int msmd_rulebook_conv_pairs(const uint32_t* out_bits, const int* out_prefix, int n_out,
                             int batch_size, const int* spatial_shape, const int* ksize,
                             const int* stride, const int* padding, const int* dilation,
                             const uint32_t* in_bits, const int* in_prefix, const int* in_perm,
                             int* out_indices /* (n_out,4) */, int* pair_fwd /* (kvol,n_out) */,
                             msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sparse convolution forward -- replaces Fsp.implicit_gemm (bug_fix/conv.py:442-447)
 * with an optional fused epilogue  y = relu?( conv * scale[c] + shift[c] + residual ).
 *   weight_krsc: the spconv-2.x parameter layout [cout, kz,ky,kx, cin]
 *                (bug_fix/conv.py:114-117); packed = [kvol, cin, cout].
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_spconv_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                            float* packed, msmd_stream_t stream);
MSMD_API int msmd_spconv_fwd(const float* features, int n_in, const float* packed_weight,
                    const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                    const float* scale, const float* shift, const float* residual, int relu,
                    float* out, msmd_stream_t stream);

/* Tensor-core path of the same contraction (tcgen05.mma kind::tf32, accumulator in TMEM,
 * fp32-level accuracy through the 3xTF32 split  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo).
 *   packed_tc: the weight re-packed once into the shared-memory image the tensor core reads
 *   (msmd_spconv_tc_packed_floats() floats, 16-byte aligned).  Supported: cout <= 256,
 *   kvol <= 32 (msmd_spconv_tc_supported); everything else uses msmd_spconv_fwd. */
MSMD_API int msmd_spconv_tc_supported(int cout, int kvol, int cin);
/* kernel variant: 0 (default) = chosen by Cout, 3 = A operand staged in tensor memory,
 * 2 = A operand in shared memory */
MSMD_API int msmd_spconv_tc_set_variant(int variant);
/* A/B switches of the host-side launch heuristics (defaults 0 = the measured rules): key 0 occupancy
 * (1 = one CTA per SM with the deepest pipeline, 2 = two CTAs per SM whenever they fit), key 1 pipeline-stage
 * cap (2..4), key 2 split-K pairs (1 = never, 2 = whenever supported), key 3 chunk blocks per pipeline stage of the
 * 16-bit kernel (2 = two: half the mbarrier round trips), key 4 the persistent split-operand kernel's weight of a
 * tile's epilogue in K-chunk times, plus one.  MSMD_TC_TUNE="occ=1,stages=3,split=1,cps=2,epi=9". */
MSMD_API int msmd_spconv_tc_set_tuning(int key, int value);
MSMD_API size_t msmd_spconv_tc_packed_floats(int cout, int kvol, int cin);
MSMD_API int msmd_spconv_tc_pack_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                        float* packed_tc, msmd_stream_t stream);
MSMD_API int msmd_spconv_fwd_tc(const float* features, int n_in, const float* packed_tc,
                                const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                const float* scale, const float* shift, const float* residual,
                                int relu, float* out, msmd_stream_t stream);
/* Same, with scratch for split-K pairs: when the tile count leaves SMs idle (<= 74 tiles) or is
 * just over one wave (149..222 tiles), two CTAs share an output tile, each takes half of the K
 * chunks, and the first to finish hands its partial sums to the second through `workspace`
 * (deterministic: a + b).  msmd_spconv_tc_workspace() returns the bytes needed, 0 = no split. */
MSMD_API size_t msmd_spconv_tc_workspace(int n_out, int cout);
MSMD_API int msmd_spconv_fwd_tc_ws(const float* features, int n_in, const float* packed_tc,
                                   const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                   const float* scale, const float* shift, const float* residual,
                                   int relu, float* out, void* workspace, size_t workspace_bytes,
                                   msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Mask-sorted tiles -- the role of mask_argsort_fwd_splits among the outputs of
 * ops.get_indice_pairs_implicit_gemm (bug_fix/conv.py:382-415): output rows are grouped by the
 * structure of their neighbour mask before the implicit GEMM tiles them, so that a 128-row tile
 * touches few kernel offsets and the convolution can skip the rest.
 *   msmd_rulebook_mask_sort   (kvol = 27 only) row_perm (n): tile slot -> output row, stable order by
 *                             a 15-bit digest of the 27-bit mask; pair_sorted (kvol, n) =
 *                             pair_fwd[:, row_perm].  Once per rulebook (a SubM rulebook serves 4-5
 *                             layers).  workspace: msmd_rulebook_mask_sort_workspace(n) bytes.
 *   msmd_spconv_fwd_tc_sorted msmd_spconv_fwd_tc_ws on the permuted table: slot s of a tile computes
 *                             output row row_perm[s] (residual read / result written there), so `out`
 *                             is in the ORIGINAL row order.  Per-row accumulation order is unchanged.
 * Opt-in (msmd_spconv_set_mask_sort / MSMD_MASK_SORT=1): correct on a B200 (profiles/r01h_quick_gpu_check.json),
 * not yet timed.
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_rulebook_mask_sort_workspace(int n);
MSMD_API int msmd_rulebook_mask_sort(const int* pair_fwd, int kvol, int n, int* row_perm,
                                     int* pair_sorted, void* workspace, size_t workspace_bytes,
                                     msmd_stream_t stream);
MSMD_API int msmd_spconv_fwd_tc_sorted(const float* features, int n_in, const float* packed_tc,
                                       const int* pair_sorted, const int* row_perm, int n_out, int cin,
                                       int cout, int kvol, const float* scale, const float* shift,
                                       const float* residual, int relu, float* out, void* workspace,
                                       size_t workspace_bytes, msmd_stream_t stream);
/* 1: the native executor (msmd_sparse_net_forward) mask-sorts every 3x3x3 SubM rulebook it builds and
 * runs the tensor-core layers that use it through msmd_spconv_fwd_tc_sorted.  Default 0. */
MSMD_API int msmd_spconv_set_mask_sort(int enable);

/* ------------------------------------------------------------------------------------
 * 16-bit operand modes of the tensor-core sparse convolution (csrc/spconv_tc16.cu) -- same contract as
 * msmd_spconv_fwd_tc (features / out fp32 in HBM, fp32 accumulate, fused epilogue), the contraction on
 * tcgen05.mma.kind::f16 with bf16 operands, 64 K elements per pipeline step:
 *   x3 = 0  "bf16":    operands rounded to bf16, one MMA per product.  The arithmetic BASELINE configs[4]
 *                      names for the train step ("bf16 sparse-conv kernels"; spconv-2.x runs fp16/bf16
 *                      features through the same Fsp.implicit_gemm call, bug_fix/conv.py:442-447).  Outside the
 *                      1e-4 inference parity bound.
 *   x3 = 1  "bf16x3":  operands split hi + lo in bf16, three MMAs per product (lo*lo dropped): ~5e-6
 *                      relative per layer, inside the parity bound, half the tensor-pipe time of 3xTF32.
 * `packed_tc16`: msmd_spconv_tc16_pack_weight image (msmd_spconv_tc16_packed_bytes bytes) of the KRSC weight
 * for the SAME x3.  `row_perm` NULL, or the slot -> row map of a mask-sorted table (msmd_rulebook_mask_sort).
 * Opt-in (MSMD_CONV_PRECISION / MSMD_TRAIN_PRECISION, msmd_conv_layer.weight_tc 2 / 3): checked on the host
 * model of tcgen05 and on a B200 (profiles/r01h_quick_gpu_check.json: correctness only, not yet timed).
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_spconv_tc16_packed_bytes(int cout, int kvol, int cin, int x3);
MSMD_API int msmd_spconv_tc16_pack_weight(const float* weight_krsc, int cout, int kvol, int cin, int x3,
                                          void* packed_tc16, msmd_stream_t stream);
MSMD_API int msmd_spconv_fwd_tc16(const float* features, int n_in, const void* packed_tc16,
                                  const int* pair_fwd, const int* row_perm, int n_out, int cin, int cout,
                                  int kvol, int x3, const float* scale, const float* shift,
                                  const float* residual, int relu, float* out, msmd_stream_t stream);
/* Kernel variant of the 16-bit modes: 2 (default) = A operand in shared memory; 3 = A operand in TENSOR memory
 * (two K elements per 32-bit column), weights-only shared memory => two CTAs per SM at Cout >= 96, plus the split-K
 * CTA pairs of msmd_spconv_fwd_tc_ws (scratch: msmd_spconv_tc16_workspace bytes, 0 = no split).  Variant 3's TMEM
 * operand layout is taken from the PTX ISA text; it is checked on the host model only (MSMD_TC16_VARIANT=3). */
MSMD_API int msmd_spconv_tc16_set_variant(int variant);
MSMD_API size_t msmd_spconv_tc16_workspace(int n_out, int cout);
MSMD_API int msmd_spconv_fwd_tc16_ws(const float* features, int n_in, const void* packed_tc16,
                                     const int* pair_fwd, const int* row_perm, int n_out, int cin, int cout,
                                     int kvol, int x3, const float* scale, const float* shift,
                                     const float* residual, int relu, float* out, void* workspace,
                                     size_t workspace_bytes, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * bf16x3 with a SPLIT-BF16 OPERAND CACHE (csrc/spconv_sb.cu) -- the same arithmetic as msmd_spconv_fwd_tc16 x3 = 1
 * (replaces the same call, Fsp.implicit_gemm at bug_fix/conv.py:442-447), with the hi / lo split of the activations
 * moved out of the gather: every activation the convolutions read also exists as a "split image"
 *       xs[row] = [ hi(0..C8) | lo(0..C8) ]  bf16,  C8 = round_up(C, 8)   (msmd_split_width(C) = 2*C8 elements a row)
 * written by the epilogue of the producing layer (`out_split`) or by msmd_split_bf16 for a network input, and the
 * gather is cp.async straight into the tensor-core operand tile.  `out` (fp32 rows) and `out_split` may each be NULL,
 * not both.  `packed_sb`: msmd_spconv_sb_pack_weight image (msmd_spconv_sb_packed_bytes bytes) of the KRSC weight.
 * msmd_conv_layer.weight_tc = 4 selects it inside msmd_sparse_net_forward (split images carved from the arena).
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_split_width(int channels);
MSMD_API int msmd_split_bf16(const float* x, int n, int channels, void* x_split, msmd_stream_t stream);
MSMD_API size_t msmd_spconv_sb_packed_bytes(int cout, int kvol, int cin);
MSMD_API int msmd_spconv_sb_pack_weight(const float* weight_krsc, int cout, int kvol, int cin, void* packed_sb,
                                        msmd_stream_t stream);
MSMD_API int msmd_spconv_fwd_sb(const void* features_split, int n_in, const void* packed_sb, const int* pair_fwd,
                                int n_out, int cin, int cout, int kvol, const float* scale, const float* shift,
                                const float* residual, int relu, float* out, void* out_split, msmd_stream_t stream);
/* Schedule of msmd_spconv_fwd_sb (A/B switch, same results): 0 = default = 2; 1 = one 128-row tile per CTA; 2 = the
 * persistent kernel: one CTA per SM slot over an equal share of the launch's (tile, K chunk) units, tiles that
 * straddle a share are summed through an L2 hand-off in fixed CTA order, epilogue overlapped with the next tile's
 * main loop.  The hand-off slots (<= 19 MB) belong to the library, one set per (device, stream), allocated on the
 * first launch on that stream. */
MSMD_API int msmd_spconv_sb_set_variant(int variant);
MSMD_API int msmd_spconv_sb_set_pdl(int enable);     /* programmatic dependent launch of the persistent kernel (default 0: measured slower end to end) */
MSMD_API int msmd_spconv_sb_uses_tile_masks(void);   /* 1 under variant 0: callers that keep rulebooks build tile masks */
/* Per-rulebook side table for the persistent schedule: bit k of tile_mask[t] (t = 128-row tile, ceil(n_out/128) words) is
 * set when some row of the tile has a pair at kernel offset k.  Built once per rulebook (spconv-2.x keeps the analogous
 * mask_argsort / pair masks next to its indice pairs, bug_fix/conv.py:382-415). */
MSMD_API int msmd_rulebook_tile_masks(const int* pair_fwd, int kvol, int n_out, unsigned* tile_mask,
                                      msmd_stream_t stream);
/* msmd_spconv_fwd_sb with the two optional tables of the persistent schedule (either may be NULL):
 *   row_perm   the pair table is a mask-sorted one (msmd_rulebook_mask_sort): tile slot i holds output row row_perm[i];
 *   tile_mask  msmd_rulebook_tile_masks of `pair_fwd`: K chunks without a pair in a tile are neither gathered nor
 *              multiplied, and the CTAs' work shares are balanced over the chunks that remain. */
MSMD_API int msmd_spconv_fwd_sb_ex(const void* features_split, int n_in, const void* packed_sb, const int* pair_fwd,
                                   const int* row_perm, const unsigned* tile_mask, int n_out, int cin, int cout,
                                   int kvol, const float* scale, const float* shift, const float* residual, int relu,
                                   float* out, void* out_split, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Sparse convolution BACKWARD (config 5, the train step) -- replaces the backward of
 * Fsp.implicit_gemm (call site bug_fix/conv.py:442-447; spconv-2.x differentiates through
 * pair_bwd / mask_argsort_bwd_splits, bug_fix/conv.py:382-415).  Arithmetic as the vendored
 * spconv-1.x indiceConvBackward (mmdet3d/ops/spconv/include/spconv/spconv_ops.h:364-457).
 *
 *   msmd_rulebook_transpose     pair_bwd (kvol, n_in): pair_bwd[k,i] = o with pair_fwd[k,o] = i,
 *                               else -1 (what get_indice_pairs_implicit_gemm returns as pair_bwd).
 *                               SubM layers do not need it: pair_bwd[k] == pair_fwd[kvol-1-k].
 *   msmd_spconv_transpose_weight  W[co,k,ci] -> Wt[ci,k',co], k' = flip_k ? kvol-1-k : k (a KRSC
 *                               weight with the channel roles swapped; feed it to
 *                               msmd_spconv_pack_weight / msmd_spconv_tc_pack_weight).
 *   msmd_spconv_bwd_data        grad_in (n_in,cin) = forward contraction of grad_out over pair_bwd
 *                               with the packed transposed weight (weight_tc: 1 = tensor-core image,
 *                               workspace as msmd_spconv_fwd_tc_ws with (n_in, cin); 0 = SIMT layout).
 *   msmd_spconv_bwd_weight      grad_weight KRSC [cout,kvol,cin] = sum over pairs of
 *                               grad_out[o,co] * features[i,ci]; deterministic (partial tiles in
 *                               `workspace`, msmd_spconv_bwd_weight_workspace() bytes, summed in a
 *                               fixed order).  Exact fp32 (FFMA).
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_rulebook_transpose(const int* pair_fwd, int kvol, int n_out, int n_in,
                                     int* pair_bwd, msmd_stream_t stream);
MSMD_API int msmd_spconv_transpose_weight(const float* weight_krsc, int cout, int kvol, int cin,
                                          int flip_k, float* weight_t, msmd_stream_t stream);
MSMD_API int msmd_spconv_bwd_data(const float* grad_out, int n_out, const float* packed_wt,
                                  int weight_tc, const int* pair_bwd, int n_in, int cin, int cout,
                                  int kvol, float* grad_in, void* workspace, size_t workspace_bytes,
                                  msmd_stream_t stream);
MSMD_API size_t msmd_spconv_bwd_weight_workspace(int n_out, int cin, int cout, int kvol);
MSMD_API int msmd_spconv_bwd_weight(const float* features, int n_in, const float* grad_out,
                                    const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                    float* grad_weight_krsc, void* workspace, size_t workspace_bytes,
                                    msmd_stream_t stream);
/* Tensor-core weight gradient (csrc/spconv_wgrad_tc.cu): same contract and workspace layout as
 * msmd_spconv_bwd_weight, the contraction over an offset's pairs on tcgen05 (3xTF32, fp32 accumulate in
 * tensor memory, deterministic slice reduction) instead of exact-fp32 FFMA.  Needs cin, cout multiples of 4,
 * cin <= 256 and 16-byte aligned features / grad_out.  msmd_spconv_set_wgrad_tc(1) (MSMD_WGRAD_TC=1) routes
 * msmd_spconv_bwd_weight / msmd_spconv_bwd_weight_workspace here whenever the shape is supported.  Opt-in:
 * checked on the host model of tcgen05 and on a B200 (profiles/r01h_quick_gpu_check.json); not yet timed. */
MSMD_API int msmd_spconv_bwd_weight_tc_supported(int cin, int cout, int kvol);
MSMD_API size_t msmd_spconv_bwd_weight_tc_workspace(int n_out, int cin, int cout, int kvol);
MSMD_API int msmd_spconv_bwd_weight_tc(const float* features, int n_in, const float* grad_out,
                                       const int* pair_fwd, int n_out, int cin, int cout, int kvol,
                                       float* grad_weight_krsc, void* workspace, size_t workspace_bytes,
                                       msmd_stream_t stream);
MSMD_API int msmd_spconv_set_wgrad_tc(int enable);
/* Backward of msmd_to_dense: rows (n,c) gathered out of a (batch, c, D, H, W) gradient. */
MSMD_API int msmd_from_dense(const int* indices, const float* dense, int n, int c, int batch_size,
                             const int* spatial_shape, float* out, msmd_stream_t stream);
/* Row of each voxel in a bit grid's ascending order (-1 = absent): the gather map that the backward
 * of Fsp.sparse_add needs (grad_a = grad_out[rows_a], grad_b = grad_out[rows_b]). */
MSMD_API int msmd_grid_rows(const int* indices, int n, int batch_size, const int* spatial_shape,
                            const uint32_t* bits, const int* prefix, int* rows, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Native executor for a chain of sparse convolutions -- the host loop of
 * SparseEncoder.forward (mmdet3d/models/middle_encoders/sparse_encoder.py:96-133) and of the
 * SparseSequential / SparseBasicBlock modules it is built from (mmdet3d/ops/sparse_block.py:
 * 103-126,161-190): ONE call runs every rulebook build and convolution of the plan.
 *   layers[i] reads activation `input` (0 = the network input, j+1 = output of layer j) and
 *   writes activation i+1;  y = relu?(conv(x) * scale + shift + acts[residual]).
 *   All intermediates (grids, rulebooks, features, indices) are carved from `arena`; on return
 *   acts[0..n_layers] describe every activation (device pointers into the arena, host counts).
 *   The only synchronisations are the N_out read-backs of the strided convolutions.
 *   MSMD_ERR_WORKSPACE: arena too small (msmd_last_error() says how far it got).
 * ---------------------------------------------------------------------------------- */
typedef struct msmd_conv_layer {
  int subm;                 /* 1 = SubMConv3d, 0 = SparseConv3d */
  int ksize[3], stride[3], padding[3], dilation[3];
  int cin, cout;
  const float* weight;      /* device: packed image selected by weight_tc */
  int weight_tc;            /* 0: msmd_spconv_pack_weight (fp32 FFMA kernel); 1: msmd_spconv_tc_pack_weight (3xTF32);
                             * 2: msmd_spconv_tc16_pack_weight x3=1 (bf16x3); 3: msmd_spconv_tc16_pack_weight x3=0 (bf16);
                             * 4: msmd_spconv_sb_pack_weight (bf16x3 through the split-bf16 operand cache) */
  const float* scale;       /* device (cout) or NULL: folded BatchNorm1d(eval) */
  const float* shift;
  int relu;
  int input;                /* activation index read by this layer */
  int residual;             /* activation index added before the ReLU, or -1 */
} msmd_conv_layer;

typedef struct msmd_sparse_desc {
  float* features;          /* device (n, channels) */
  int* indices;             /* device (n, 4) (b,z,y,x) */
  int n, channels;
  int spatial_shape[3];
} msmd_sparse_desc;

MSMD_API int msmd_sparse_net_forward(const msmd_conv_layer* layers, int n_layers,
                                     const float* features, const int* indices, int n, int channels,
                                     int batch_size, const int* spatial_shape /* host [3] */,
                                     void* arena, size_t arena_bytes,
                                     msmd_sparse_desc* acts /* host [n_layers + 1] */,
                                     msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Gating + concatenation of one Gated Modality-Aware stage (csrc/gma.cu) -- replaces, for one sample per GPU,
 * sparse_multimodal_encoder_painting.py:371-377 (cross gate of the only-2D voxels), :391-401 (gate of the mixed
 * voxels) and :414-425 (zero-padded concatenation into the unified voxel list) with ONE kernel.
 *   unified rows: [0, n_only3)            [ y_only3 | 0 x 64 ]                      coords idx_only3
 *                 next max(n_only2, 1)     [ 0 x c3 | relu(W_cross g + b) * feat2[only2_rows[j]] ],  g = feat3[nn_idx[j]]
 *                                          or the dummy embedding when nn_idx[j] < 0; coords only2_bzyx
 *                 next max(n_mix, 1)       [ feat3[syn3[p]] | relu(W_gate feat3[syn3[p]] + b) * feat2[syn2[p]] ], coords
 *                                          bz2[syn2[p]]
 *   An empty group (only2_rows == NULL, or n_mix == 0) contributes the one all-zero voxel pad_missing_batch_id appends
 *   (:208-225).  w_* are nn.Linear weights [64][c3], c2 must be 64.
 * msmd_gather_rows: out[i] = features[rows[i]] (+ the (b,z,y,x) rows when coords4 is given).
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_gather_rows(const float* features, int channels, const int* coords4, const long long* rows, int n,
                              float* out_features, int* out_coords4, msmd_stream_t stream);
MSMD_API int msmd_gma_assemble(const float* y_only3, const int* idx_only3, int n_only3, const float* feat3, int n3,
                               int c3, const float* feat2, const int* bz2, int n2, int c2,
                               const long long* only2_rows, const int* only2_bzyx, const long long* nn_idx,
                               int n_only2, const long long* syn3, const long long* syn2, int n_mix,
                               const float* dummy, const float* w_cross, const float* b_cross, const float* w_gate,
                               const float* b_gate, float* unified_features, int* unified_indices,
                               msmd_stream_t stream);   /* unified_features or unified_indices may be NULL (not both) */

/* The executor's geometry stream (a cudaStream_t) of the current device.  Index sets and rulebooks of a
 * msmd_sparse_net_forward call are complete once everything queued on it at the call's return has run: consumers of
 * the COORDINATES only (MSMDFusion.py:251-325 voxel_modality_split, :276-323 fps_NN_fast) may wait for an event
 * recorded there instead of for the feature convolutions on the caller's stream. */
MSMD_API int msmd_executor_geometry_stream(void** stream_out);

/* Same, reporting how many bytes of the arena the call carved (the caller may keep bump-allocating behind them).
 * flags: MSMD_NET_INDICES_ON_GEOMETRY_STREAM = `indices` were produced on msmd_executor_geometry_stream() (not on
 * `stream`): the call's rulebook chain does not wait for `stream` and may run ahead of the convolutions queued there. */
#define MSMD_NET_INDICES_ON_GEOMETRY_STREAM 1
MSMD_API int msmd_sparse_net_forward_ex(const msmd_conv_layer* layers, int n_layers, const float* features,
                                        const int* indices, int n, int channels, int batch_size,
                                        const int* spatial_shape, void* arena, size_t arena_bytes,
                                        msmd_sparse_desc* acts, size_t* arena_used, int flags, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * One whole stage of the Gated Modality-Aware convolution in ONE call (csrc/gma.cu) -- the body of
 * SparseMultiModalEncoderPaint.forward's loop (sparse_multimodal_encoder_painting.py:433-459) for one sample per GPU:
 *   only-3D rows gathered -> grouped_sp_conv_blocks_3D chain -> gates + concatenation (msmd_gma_assemble)
 *   -> aggregation_blocks chain -> Fsp.sparse_add with the previous stage's output (when given) -> downscale chain.
 * The three chains are msmd_conv_layer lists as for msmd_sparse_net_forward; all intermediates and the result are
 * carved from `arena` (MSMD_ERR_WORKSPACE: too small); `out` describes the stage output.  Everything that depends on
 * COORDINATES only (row-list gathers of coordinates, the unified coordinate list, bit grids, rulebooks, the union of
 * sparse_add and both N_out read-backs) runs on msmd_executor_geometry_stream(), after `indices_ready_event` (the
 * event behind which only3_rows / only2_* / nn_idx / syn* / bz2 are complete; NULL: after `stream`); the gathers of
 * features, the gate kernel and the convolutions run on `stream`.  The host blocks on the geometry stream only, so
 * the rulebooks of a stage -- and of the next stages -- are built while earlier convolutions still run.
 * ---------------------------------------------------------------------------------- */
typedef struct msmd_gma_stage {
  const msmd_conv_layer* only3d; int n_only3d;
  const msmd_conv_layer* agg;    int n_agg;
  const msmd_conv_layer* down;   int n_down;
  const float* w_cross; const float* b_cross;   /* cross_gate_control[stage][0]: nn.Linear(c3 -> 64) */
  const float* w_gate;  const float* b_gate;    /* gate_control[stage][0] */
  int c3, c2;
} msmd_gma_stage;

MSMD_API int msmd_gma_stage_forward(const msmd_gma_stage* stage, const float* feat3, const int* bz3, int n3,
                                    const float* feat2, const int* bz2, int n2, const long long* only3_rows,
                                    int n_only3, const long long* only2_rows, const int* only2_bzyx,
                                    const long long* nn_idx, int n_only2, const long long* syn3,
                                    const long long* syn2, int n_mix, const float* dummy, const float* prev_features,
                                    const int* prev_indices, int n_prev, int batch_size,
                                    const int* spatial_shape /* host [3] */, void* arena, size_t arena_bytes,
                                    msmd_sparse_desc* out, void* indices_ready_event /* cudaEvent_t or NULL */,
                                    msmd_stream_t stream);

/* SparseConvTensor.dense(): (n,c) rows -> (batch, c, D, H, W), zero-filled inside.
 * spconv-1.x equivalent mmdet3d/ops/spconv/structure.py:54-66. */
MSMD_API int msmd_to_dense(const int* indices, const float* features, int n, int c, int batch_size,
                  const int* spatial_shape, float* out, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Furthest point sampling -- replaces furthest_point_sample(xyz (B,N,3) f32, m) -> (B,m) i32
 *   mmdet3d/ops/furthest_point_sample/furthest_point_sample.py:8-37, CUDA kernel
 *   src/furthest_point_sample_cuda.cu:25-140 (start index 0, arg-max tie-break reproduced).
 *   One batch element per call: xyz (n,3), idx (m).
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_fps_workspace(int n);
/* A/B switch: CTA width of the thread-block-cluster FPS kernel (0 = default heuristic | 256 | 512 | 1024) */
MSMD_API int msmd_fps_set_threads(int threads);
MSMD_API int msmd_fps(const float* xyz, int n, int m, int* idx, void* workspace,
                      size_t workspace_bytes, msmd_stream_t stream);

/* Ball query -- replaces ball_query(min_radius, max_radius, nsample, xyz, center_xyz)
 *   mmdet3d/ops/ball_query/ball_query.py:8-46, src/ball_query_cuda.cu:11-55.
 *   xyz (n,3) candidates, centers (m,3); idx (m,nsample) is zero-filled inside. */
MSMD_API int msmd_ball_query(const float* xyz, int n, const float* centers, int m,
                             float min_radius, float max_radius, int nsample, int* idx,
                             msmd_stream_t stream);

/* Nearest key voxel per query voxel on integer (z,y,x) coordinates; rows are
 * `*_stride` ints apart and point at the z column.  val = ||q-key||_2 (fp32), idx = first
 * minimal key.  Replaces torch.norm(...).min(-1) of
 * sparse_multimodal_encoder_painting.py:289-291 / :302-305 without the (Q,N,3) temporary. */
MSMD_API int msmd_nn_search(const int* query, int query_stride, int nq, const int* key,
                            int key_stride, int nk, float* val, int* idx, msmd_stream_t stream);

/* query_NN_key_idx[group] = nn  (painting.py:311-321).  group (m,nsample) from ball query
 * (NULL: every query is its own representative, the Q <= fps_num branch :288-293);
 * out (nq) int64 = nn_idx + base or -1.  Duplicate targets: the last (representative,
 * slot) pair in row-major order wins.  winner_scratch: nq ints. */
MSMD_API int msmd_group_assign(const int* group, int m, int nsample, const float* val,
                               const int* nn_idx, float dist_thresh, int nq, int base,
                               int* winner_scratch, long long* out, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * voxel_modality_split for ONE sample -- replaces MSMDFusion.py:251-325 (+ type_assign
 * :27-45): float32 key z*1e6+y*1e3+x, stable sort, one-to-one pairing of equal keys.
 *   coord3 (n3,4) / coord2 (n2,4) i32 (b,z,y,x) rows of this sample
 *   mix3 (n3) / mix2 (n2) i32: 1 = voxel exists in both modalities
 *   syn3 / syn2: int64 row ids (+offset) of the paired voxels in sorted-key order; capacity
 *   min(n3,n2); num_mix (device, TWO ints): [0] = number of pairs, [1] = overflow flag
 * msmd_modality_split: hash table on the float key (counts + up to 8 row ids per key and set) + a shared-memory sort of
 * the paired rows only -- 5 launches; sets num_mix[1] = 1 when a key run exceeds 8 rows or there are more than 4096
 * pairs: the caller then runs msmd_modality_split_sort (two stable radix sorts + binary-search merge, any size; it
 * writes num_mix[0] only).  Both produce the reference's result bit for bit.
 * ---------------------------------------------------------------------------------- */
MSMD_API size_t msmd_modality_split_workspace(int n3, int n2);
MSMD_API int msmd_modality_split_sort(const int* coord3, int n3, const int* coord2, int n2, long long offset3,
                                      long long offset2, int* mix3, int* mix2, long long* syn3, long long* syn2,
                                      int* num_mix, void* workspace, size_t workspace_bytes, msmd_stream_t stream);
MSMD_API int msmd_modality_split(const int* coord3, int n3, const int* coord2, int n2,
                                 long long offset3, long long offset2, int* mix3, int* mix2,
                                 long long* syn3, long long* syn2, int* num_mix, void* workspace,
                                 size_t workspace_bytes, msmd_stream_t stream);

/* Rows i (ascending) with flags[i] == 0 -> out_rows (capacity n, int64), *count (device).
 * Replaces the boolean-mask indexing `indices[:, 1] == 0` of
 * sparse_multimodal_encoder_painting.py:338-343 (only-3D / only-2D voxel selection) without
 * the host synchronisation of torch.nonzero: the caller already knows the count
 * (n - number of mixed pairs).  workspace: msmd_scan_workspace() bytes. */
MSMD_API int msmd_compact_unflagged(const int* flags, int n, long long* out_rows, int* count,
                                    void* workspace, size_t workspace_bytes, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Fsp.sparse_add(a, b) -- call site sparse_multimodal_encoder_painting.py:455.
 *   step 1 builds the union bit grid (bits/prefix sized by msmd_grid_num_words) and
 *   *num_out; step 2 (needs num_out on the host) writes out_indices (ascending linear
 *   order) and out_features = sum of coincident rows.  The grid is reusable as the index
 *   structure of the output tensor.
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_sparse_add_outputs(const int* idx_a, int na, const int* idx_b, int nb,
                                     int batch_size, const int* spatial_shape, uint32_t* bits,
                                     int* prefix, int* num_out, void* workspace,
                                     size_t workspace_bytes, msmd_stream_t stream);
MSMD_API int msmd_sparse_add_finish(const uint32_t* bits, const int* prefix, int n_out,
                                    const int* idx_a, const float* feat_a, int na,
                                    const int* idx_b, const float* feat_b, int nb, int c,
                                    int batch_size, const int* spatial_shape, int* out_indices,
                                    float* out_features, msmd_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Virtual-point lift -- replaces the per-camera loop of get_foreground2D
 * (MSMDFusion.py:189-230): out[p] = [ points[p] (point_dims) | feat[cam,:,v,u] * score ],
 * score = ReLU(w . [feat, depth, lidar2img[cam] (16)] + b).  img_feat is addressed by
 * element strides (any layout); pixels (M,3) = (u,v,depth) in network-input pixels.
 * ---------------------------------------------------------------------------------- */
MSMD_API int msmd_lift_gather(const float* img_feat, long long stride_cam, long long stride_c,
                              long long stride_y, long long stride_x, int channels, int height,
                              int width, const float* pixels, const int* cam_ids,
                              const float* points, int point_dims, int num_points,
                              const float* lidar2img, float downscale, const float* score_weight,
                              float score_bias, float* out, msmd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSMD_B200_H_ */
