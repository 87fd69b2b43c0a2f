/*
 * msmd_oracle.c -- CPU restatement of the MSMDFusion voxel-space fusion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under msmdfusion_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * Every function restates a reference algorithm and cites the reference file:line
 * (paths relative to /root/reference).  The sparse-convolution arithmetic lives in
 * the un-vendored third-party package spconv v2.1.21 (README.md:19-20); for those
 * functions the published algorithm is restated and anchored on the reference call
 * sites (bug_fix/conv.py:382-396, :442-447) and on the in-tree spconv-1.x geometry
 * (mmdet3d/ops/spconv/include/spconv/geometry.h:24-85).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (see oracle/build.py).  No -ffast-math: the
 * voxel coordinate arithmetic must stay IEEE float32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* small open-addressing hash: int64 key -> int32 value                       */
/* ------------------------------------------------------------------------- */
typedef struct {
  int64_t *keys;
  int32_t *vals;
  uint64_t mask;
} orc_map;

static uint64_t orc_hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33; return x;
}

static int orc_map_init(orc_map *m, int64_t n) {
  uint64_t cap = 16;
  while (cap < (uint64_t)(2 * n + 1)) cap <<= 1;
  m->keys = (int64_t *)malloc(cap * sizeof(int64_t));
  m->vals = (int32_t *)malloc(cap * sizeof(int32_t));
  if (!m->keys || !m->vals) return -1;
  for (uint64_t i = 0; i < cap; ++i) m->keys[i] = -1;
  m->mask = cap - 1;
  return 0;
}
static void orc_map_free(orc_map *m) { free(m->keys); free(m->vals); }

/* returns pointer to value slot; *fresh = 1 when the key was inserted now */
static int32_t *orc_map_get(orc_map *m, int64_t key, int insert, int *fresh) {
  uint64_t h = orc_hash64((uint64_t)key) & m->mask;
  for (;;) {
    if (m->keys[h] == key) { if (fresh) *fresh = 0; return &m->vals[h]; }
    if (m->keys[h] == -1) {
      if (!insert) return NULL;
      m->keys[h] = key; if (fresh) *fresh = 1; return &m->vals[h];
    }
    h = (h + 1) & m->mask;
  }
}

/* ------------------------------------------------------------------------- */
/* hard_voxelize -- mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-142          */
/* ------------------------------------------------------------------------- */
/* grid_size[i] = round((range[3+i]-range[i]) / voxel_size[i])  (:121-124)     */
void orc_grid_size(const float *voxel_size, const float *coors_range, int *grid_size) {
  for (int i = 0; i < 3; ++i)
    grid_size[i] = (int)roundf((coors_range[3 + i] - coors_range[i]) / voxel_size[i]);
}

/*
 * points (N,C) f32; outputs must be zero-filled by the caller exactly as
 * mmdet3d/ops/voxel/voxelize.py:45-50 does: voxels (max_voxels,max_points,C),
 * coors (max_voxels,3) in (z,y,x) order, num_points_per_voxel (max_voxels,).
 * Returns voxel_num.  Serial first-appearance order with the `break` at
 * max_voxels (voxelization_cpu.cpp:68-96).
 */
int orc_hard_voxelize(const float *points, int N, int C, const float *voxel_size,
                      const float *coors_range, int max_points, int max_voxels,
                      float *voxels, int *coors, int *num_points_per_voxel) {
  int grid[3];
  orc_grid_size(voxel_size, coors_range, grid);
  orc_map map;
  if (orc_map_init(&map, N) != 0) return -1;
  int voxel_num = 0;
  for (int i = 0; i < N; ++i) {
    /* dynamic_voxelize_kernel, voxelization_cpu.cpp:20-37 */
    int c[3];
    int failed = 0;
    for (int j = 0; j < 3; ++j) {
      float q = (points[(size_t)i * C + j] - coors_range[j]) / voxel_size[j];
      int cj = (int)floorf(q);
      if (cj < 0 || cj >= grid[j]) { failed = 1; break; }
      c[2 - j] = cj; /* reversed: (z,y,x) */
    }
    if (failed) continue; /* coor[i][0] == -1 -> continue (:71) */
    int64_t key = ((int64_t)c[0] * grid[1] + c[1]) * grid[0] + c[2];
    int fresh = 0;
    int32_t *slot = orc_map_get(&map, key, 1, &fresh);
    int voxelidx;
    if (fresh) {
      if (max_voxels != -1 && voxel_num >= max_voxels) break; /* :78 */
      voxelidx = voxel_num++;
      *slot = voxelidx;
      for (int k = 0; k < 3; ++k) coors[(size_t)voxelidx * 3 + k] = c[k];
    } else {
      voxelidx = *slot;
    }
    int num = num_points_per_voxel[voxelidx];
    if (max_points == -1 || num < max_points) { /* :90-95 */
      memcpy(voxels + ((size_t)voxelidx * max_points + num) * C, points + (size_t)i * C,
             sizeof(float) * C);
      num_points_per_voxel[voxelidx] = num + 1;
    }
  }
  orc_map_free(&map);
  return voxel_num;
}

/* ------------------------------------------------------------------------- */
/* sparse-conv geometry (spconv v2.1.21 semantics; see header comment)         */
/* ------------------------------------------------------------------------- */
/* spconv.ops.get_conv_output_size: (in + 2p - d(k-1) - 1)/s + 1; same formula as
 * mmdet3d/ops/spconv/ops.py:20-31 */
void orc_conv_out_shape(const int *shape, const int *ksize, const int *stride,
                        const int *pad, const int *dil, int *out_shape) {
  for (int i = 0; i < 3; ++i) {
    int v = (shape[i] + 2 * pad[i] - dil[i] * (ksize[i] - 1) - 1) / stride[i] + 1;
    out_shape[i] = v;
  }
}

static int64_t lin_index(int b, int z, int y, int x, const int *shape) {
  return (((int64_t)b * shape[0] + z) * shape[1] + y) * shape[2] + x;
}

/*
 * Submanifold rulebook.  Output voxel o == input voxel o (indices unchanged);
 * pair_fwd[k*N + o] = row of the active input at  coord(o) + (k_d - ksize_d/2)*dil_d
 * or -1.  Kernel offset k = (kz*KY + ky)*KX + kx (kx fastest), cross-correlation,
 * as geometry.h:24-85 (getValidOutPos) enumerates offsets.
 * If several input rows share one coordinate the LARGEST row wins (the reference
 * hash insert leaves the winner undefined).
 */
int orc_subm_rulebook(const int *indices, int N, const int *shape, const int *ksize,
                      const int *dil, int *pair_fwd) {
  orc_map map;
  if (orc_map_init(&map, N) != 0) return -1;
  for (int i = 0; i < N; ++i) {
    const int *c = indices + (size_t)i * 4;
    int32_t *s = orc_map_get(&map, lin_index(c[0], c[1], c[2], c[3], shape), 1, NULL);
    *s = i;
  }
  int K = ksize[0] * ksize[1] * ksize[2];
  for (int k = 0; k < K; ++k) {
    int kz = k / (ksize[1] * ksize[2]), ky = (k / ksize[2]) % ksize[1], kx = k % ksize[2];
    int dz = (kz - ksize[0] / 2) * dil[0], dy = (ky - ksize[1] / 2) * dil[1],
        dx = (kx - ksize[2] / 2) * dil[2];
    for (int o = 0; o < N; ++o) {
      const int *c = indices + (size_t)o * 4;
      int z = c[1] + dz, y = c[2] + dy, x = c[3] + dx;
      int v = -1;
      if (z >= 0 && z < shape[0] && y >= 0 && y < shape[1] && x >= 0 && x < shape[2]) {
        int32_t *s = orc_map_get(&map, lin_index(c[0], z, y, x, shape), 0, NULL);
        if (s) v = *s;
      }
      pair_fwd[(size_t)k * N + o] = v;
    }
  }
  orc_map_free(&map);
  return 0;
}

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

/*
 * Regular (strided) sparse conv rulebook.  Output set = { o : exists active i, k
 * with o*s - p + k*d == i }, ordered ASCENDING by linear index
 * ((b*D+z)*H+y)*W+x over the OUTPUT shape (spconv-2.x generate_conv_inds: sort +
 * unique of candidate output ids; same order as the spconv-1.x GPU path,
 * mmdet3d/ops/spconv/include/spconv/spconv_ops.h:129-136).
 * out_indices capacity: caller provides N*K rows worst case (or any bound >= N_out).
 * pair_fwd is written as [K, out_capacity_stride] with stride = *n_out (dense),
 * so it is produced in a second call: pass pair_fwd == NULL to only count.
 * Returns N_out (or -1).
 */
int orc_conv_rulebook(const int *indices, int N, const int *shape, const int *ksize,
                      const int *stride, const int *pad, const int *dil,
                      int *out_indices, int out_capacity, int *pair_fwd) {
  int oshape[3];
  orc_conv_out_shape(shape, ksize, stride, pad, dil, oshape);
  int K = ksize[0] * ksize[1] * ksize[2];
  int64_t *cand = (int64_t *)malloc(sizeof(int64_t) * (size_t)N * K + 8);
  if (!cand) return -1;
  size_t nc = 0;
  for (int i = 0; i < N; ++i) {
    const int *c = indices + (size_t)i * 4;
    for (int k = 0; k < K; ++k) {
      int kk[3] = {k / (ksize[1] * ksize[2]), (k / ksize[2]) % ksize[1], k % ksize[2]};
      int o[3];
      int ok = 1;
      for (int d = 0; d < 3; ++d) {
        int t = c[1 + d] + pad[d] - kk[d] * dil[d];
        if (t < 0 || t % stride[d] != 0) { ok = 0; break; }
        o[d] = t / stride[d];
        if (o[d] >= oshape[d]) { ok = 0; break; }
      }
      if (ok) cand[nc++] = lin_index(c[0], o[0], o[1], o[2], oshape);
    }
  }
  qsort(cand, nc, sizeof(int64_t), cmp_i64);
  size_t nu = 0;
  for (size_t i = 0; i < nc; ++i)
    if (i == 0 || cand[i] != cand[i - 1]) cand[nu++] = cand[i];
  if ((int)nu > out_capacity) { free(cand); return -2; }
  for (size_t o = 0; o < nu; ++o) {
    int64_t L = cand[o];
    int x = (int)(L % oshape[2]); L /= oshape[2];
    int y = (int)(L % oshape[1]); L /= oshape[1];
    int z = (int)(L % oshape[0]); L /= oshape[0];
    out_indices[o * 4 + 0] = (int)L;
    out_indices[o * 4 + 1] = z;
    out_indices[o * 4 + 2] = y;
    out_indices[o * 4 + 3] = x;
  }
  free(cand);
  if (pair_fwd) {
    orc_map map;
    if (orc_map_init(&map, N) != 0) return -1;
    for (int i = 0; i < N; ++i) {
      const int *c = indices + (size_t)i * 4;
      *orc_map_get(&map, lin_index(c[0], c[1], c[2], c[3], shape), 1, NULL) = i;
    }
    for (int k = 0; k < K; ++k) {
      int kk[3] = {k / (ksize[1] * ksize[2]), (k / ksize[2]) % ksize[1], k % ksize[2]};
      for (size_t o = 0; o < nu; ++o) {
        const int *c = out_indices + o * 4;
        int in[3];
        int ok = 1;
        for (int d = 0; d < 3; ++d) {
          in[d] = c[1 + d] * stride[d] - pad[d] + kk[d] * dil[d];
          if (in[d] < 0 || in[d] >= shape[d]) { ok = 0; break; }
        }
        int v = -1;
        if (ok) {
          int32_t *s = orc_map_get(&map, lin_index(c[0], in[0], in[1], in[2], shape), 0, NULL);
          if (s) v = *s;
        }
        pair_fwd[(size_t)k * nu + o] = v;
      }
    }
    orc_map_free(&map);
  }
  return (int)nu;
}

/*
 * Sparse conv forward: out[o,co] = sum_k sum_ci x[pair_fwd[k,o], ci] * W[co,k,ci]
 * W is the spconv-2.x KRSC parameter layout [Cout, kz,ky,kx, Cin]
 * (bug_fix/conv.py:114-117); fp32 accumulate; bias=False on the path
 * (mmdet3d/ops/sparse_block.py:176).  Conv math as spconv_ops.h:299-356
 * (gather -> GEMM -> scatter-add per kernel offset).
 */
int orc_spconv_fwd(const float *feat, const float *weight, const int *pair_fwd, int n_out,
                   int cin, int cout, int K, float *out) {
  /* repack W to [K][Cin][Cout] so the inner loop vectorises over co */
  float *wt = (float *)malloc(sizeof(float) * (size_t)K * cin * cout);
  if (!wt) return -1;
  for (int co = 0; co < cout; ++co)
    for (int k = 0; k < K; ++k)
      for (int ci = 0; ci < cin; ++ci)
        wt[((size_t)k * cin + ci) * cout + co] = weight[((size_t)co * K + k) * cin + ci];
#pragma omp parallel for schedule(static)
  for (int o = 0; o < n_out; ++o) {
    float *acc = out + (size_t)o * cout;
    for (int co = 0; co < cout; ++co) acc[co] = 0.f;
    for (int k = 0; k < K; ++k) {
      int p = pair_fwd[(size_t)k * n_out + o];
      if (p < 0) continue;
      const float *x = feat + (size_t)p * cin;
      const float *wk = wt + (size_t)k * cin * cout;
      for (int ci = 0; ci < cin; ++ci) {
        float xv = x[ci];
        const float *wr = wk + (size_t)ci * cout;
        for (int co = 0; co < cout; ++co) acc[co] += xv * wr[co];
      }
    }
  }
  free(wt);
  return 0;
}

/*
 * Sparse conv backward (config 5, the train step).  The reference path calls spconv-2.x
 * implicit_gemm backward through pair_bwd (bug_fix/conv.py:442-447); the arithmetic is the one
 * the vendored spconv-1.x spells out (spconv_ops.h:364-457): per kernel offset k, with the
 * offset's pairs (i -> o):   dW_k = X_k^T dY_k      dX_k = dY_k W_k^T   (scatter-added into dX).
 *   grad_in[i,ci]  = sum_{k, o: pair_fwd[k,o]=i} sum_co grad_out[o,co] * W[co,k,ci]
 *   grad_w[co,k,ci] = sum_{o: pair_fwd[k,o]>=0}  grad_out[o,co] * x[pair_fwd[k,o],ci]
 * grad_w is KRSC like the parameter; it is accumulated in double (a checker should carry less
 * rounding than the thing it checks), grad_in in fp32 like the forward.
 */
int orc_spconv_bwd(const float *feat, const float *weight, const int *pair_fwd,
                   const float *grad_out, int n_in, int n_out, int cin, int cout, int K,
                   float *grad_in, float *grad_w) {
  if (grad_in) {
    for (size_t t = 0; t < (size_t)n_in * cin; ++t) grad_in[t] = 0.f;
    for (int o = 0; o < n_out; ++o) {
      const float *dy = grad_out + (size_t)o * cout;
      for (int k = 0; k < K; ++k) {
        int p = pair_fwd[(size_t)k * n_out + o];
        if (p < 0) continue;
        float *dx = grad_in + (size_t)p * cin;
        for (int co = 0; co < cout; ++co) {
          float g = dy[co];
          const float *wr = weight + ((size_t)co * K + k) * cin;
          for (int ci = 0; ci < cin; ++ci) dx[ci] += g * wr[ci];
        }
      }
    }
  }
  if (grad_w) {
    int fail = 0;
#pragma omp parallel for schedule(dynamic)
    for (int k = 0; k < K; ++k) {
      double *acc = (double *)calloc((size_t)cout * cin, sizeof(double));
      if (!acc) { fail = 1; continue; }
      for (int o = 0; o < n_out; ++o) {
        int p = pair_fwd[(size_t)k * n_out + o];
        if (p < 0) continue;
        const float *x = feat + (size_t)p * cin;
        const float *dy = grad_out + (size_t)o * cout;
        for (int co = 0; co < cout; ++co) {
          double g = dy[co];
          double *ar = acc + (size_t)co * cin;
          for (int ci = 0; ci < cin; ++ci) ar[ci] += g * (double)x[ci];
        }
      }
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          grad_w[((size_t)co * K + k) * cin + ci] = (float)acc[(size_t)co * cin + ci];
      free(acc);
    }
    if (fail) return -1;
  }
  return 0;
}

/* pair_bwd (K, n_in): pair_bwd[k,i] = o with pair_fwd[k,o] = i, else -1 (at most one such o per
 * (k,i): o = (i + pad - k*dil) / stride).  What spconv-2.x returns next to pair_fwd
 * (bug_fix/conv.py:382-415). */
void orc_pair_transpose(const int *pair_fwd, int K, int n_out, int n_in, int *pair_bwd) {
  for (size_t t = 0; t < (size_t)K * n_in; ++t) pair_bwd[t] = -1;
  for (int k = 0; k < K; ++k)
    for (int o = 0; o < n_out; ++o) {
      int p = pair_fwd[(size_t)k * n_out + o];
      if (p >= 0 && p < n_in) pair_bwd[(size_t)k * n_in + p] = o;
    }
}

/* ------------------------------------------------------------------------- */
/* furthest point sampling                                                     */
/* mmdet3d/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:11-140  */
/* ------------------------------------------------------------------------- */
/* opt_n_threads (:11-15): truncating log2 computed in double */
int orc_fps_block(int n) {
  int pow_2 = (int)(log((double)n) / log(2.0));
  int t = 1 << pow_2;
  if (t > 1024) t = 1024;
  if (t < 1) t = 1;
  return t;
}

/*
 * xyz (n,3) f32, temp (n) initialised by the caller to 1e10
 * (furthest_point_sample.py:28), idx (m) out.  Emulates the block-wide arg-max
 * including its tie-break: each thread keeps the FIRST maximal k of its strided
 * set (strict >, :69-70); the shared-memory tree keeps the lower slot on ties
 * (__update, :17-23).
 */
int orc_fps(const float *xyz, int n, int m, float *temp, int *idx) {
  if (m <= 0) return 0;
  int block = orc_fps_block(n);
  float *dists = (float *)malloc(sizeof(float) * block);
  int *dists_i = (int *)malloc(sizeof(int) * block);
  if (!dists || !dists_i) return -1;
  int old = 0;
  idx[0] = 0;
  for (int j = 1; j < m; ++j) {
    float x1 = xyz[old * 3 + 0], y1 = xyz[old * 3 + 1], z1 = xyz[old * 3 + 2];
    for (int tid = 0; tid < block; ++tid) {
      int besti = 0;
      float best = -1.f;
      for (int k = tid; k < n; k += block) {
        float x2 = xyz[k * 3 + 0], y2 = xyz[k * 3 + 1], z2 = xyz[k * 3 + 2];
        float d = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
        float d2 = d < temp[k] ? d : temp[k];
        temp[k] = d2;
        if (d2 > best) { besti = k; best = d2; }
      }
      dists[tid] = best;
      dists_i[tid] = besti;
    }
    for (int s = block / 2; s >= 1; s >>= 1) {
      for (int tid = 0; tid < s; ++tid) {
        float v1 = dists[tid], v2 = dists[tid + s];
        int i1 = dists_i[tid], i2 = dists_i[tid + s];
        dists[tid] = v1 > v2 ? v1 : v2;
        dists_i[tid] = v2 > v1 ? i2 : i1;
      }
    }
    old = dists_i[0];
    idx[j] = old;
  }
  free(dists);
  free(dists_i);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* ball query -- mmdet3d/ops/ball_query/src/ball_query_cuda.cu:11-55           */
/* ------------------------------------------------------------------------- */
/* xyz (n,3) candidates, new_xyz (m,3) centres, idx (m,nsample) zero-initialised by
 * the caller (ball_query.py:35). */
int orc_ball_query(const float *xyz, int n, const float *new_xyz, int m, float min_radius,
                   float max_radius, int nsample, int *idx) {
  float max_r2 = max_radius * max_radius, min_r2 = min_radius * min_radius;
#pragma omp parallel for schedule(static)
  for (int p = 0; p < m; ++p) {
    float nx = new_xyz[p * 3 + 0], ny = new_xyz[p * 3 + 1], nz = new_xyz[p * 3 + 2];
    int *row = idx + (size_t)p * nsample;
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
      float x = xyz[k * 3 + 0], y = xyz[k * 3 + 1], z = xyz[k * 3 + 2];
      float d2 = (nx - x) * (nx - x) + (ny - y) * (ny - y) + (nz - z) * (nz - z);
      if (d2 == 0 || (d2 >= min_r2 && d2 < max_r2)) {
        if (cnt == 0)
          for (int l = 0; l < nsample; ++l) row[l] = k;
        row[cnt] = k;
        ++cnt;
        if (cnt >= nsample) break;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* brute-force nearest key                                                     */
/* sparse_multimodal_encoder_painting.py:289-291, :302-305                     */
/* dist = ||q - key||_2 (fp32), (val, idx) = min over keys, first minimum wins  */
/* ------------------------------------------------------------------------- */
int orc_nn_search(const int *query, int nq, const int *key, int nk, float *val, int *idx) {
#pragma omp parallel for schedule(static)
  for (int q = 0; q < nq; ++q) {
    float best = INFINITY;
    int besti = 0;
    for (int k = 0; k < nk; ++k) {
      float dz = (float)query[q * 3 + 0] - (float)key[k * 3 + 0];
      float dy = (float)query[q * 3 + 1] - (float)key[k * 3 + 1];
      float dx = (float)query[q * 3 + 2] - (float)key[k * 3 + 2];
      float d = sqrtf(dz * dz + dy * dy + dx * dx);
      if (d < best) { best = d; besti = k; }
    }
    val[q] = best;
    idx[q] = besti;
  }
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
