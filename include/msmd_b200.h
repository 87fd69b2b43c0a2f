/*
 * msmd_b200.h -- C ABI of libmsmd_b200.so: the B200 (sm_100a) implementation of the
 * MSMDFusion voxel-space fusion hot path.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless a comment
 *     says "host".  No torch types cross this boundary.
 *   - the library never allocates: outputs and scratch ("workspace") are provided by the
 *     caller; every op with scratch has a *_workspace() size query.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *     synchronises.  Counts produced on the device (voxel_num, N_out) are written to a
 *     device int the caller reads back when it needs the value on the host.
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

/* SparseConvTensor.dense(): (n,c) rows -> (batch, c, D, H, W), zero-filled inside.
 * spconv-1.x equivalent mmdet3d/ops/spconv/structure.py:54-66. */
MSMD_API int msmd_to_dense(const int* indices, const float* features, int n, int c, int batch_size,
                  const int* spatial_shape, float* out, msmd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSMD_B200_H_ */
