"""TEST INFRASTRUCTURE.  Runs the reference's own `fps_NN_fast` glue where it lies.

`SparseMultiModalEncoderPaint.fps_NN_fast` (mmdet3d/models/middle_encoders/
sparse_multimodal_encoder_painting.py:276-323) is pure torch around two CUDA-only ops of the
reference (`furthest_point_sample`, `ball_query`).  Its method body is compiled here from the
reference's own source text, read at run time from /root/reference (nothing is copied into this
repository), with those two names bound to the C restatements in oracle/c/msmd_oracle.c -- which are
themselves pinned against the reference's known-answer vectors (tests/test_oracle.py).  Everything
else -- the torch.norm distance matrix, the first-minimum tie-break of `min(-1)`, the distance
threshold, the repeat / mask / duplicate-index scatter -- is the reference's code executed by torch
on the CPU.  The final `index_put_` has duplicate indices (a query inside several balls): its winner is
undefined in the reference (CUDA).  torch's CPU kernel is sequential -- last write wins -- when run with
one intra-op thread, so `fps_nn_fast` below pins torch to one thread for the call: that is the one
deterministic execution of the reference's code, and the convention oracle/cpu.py:fps_nn_fast and the
CUDA path (msmd_group_assign) fix.

Used by tests/test_oracle.py and tests/golden/make_golden_assign.py.  Needs /root/reference.
"""
import numpy as np

from . import cpu
from .ref_inplace import available, load_def  # noqa: F401

REF_ENCODER = 'mmdet3d/models/middle_encoders/sparse_multimodal_encoder_painting.py'
_fn = None


def _fps(xyz, m):
    """(1,N,3) float tensor -> (1,m) int32 tensor, like ops/furthest_point_sample (CUDA-only there)."""
    import torch
    return torch.from_numpy(cpu.furthest_point_sample(xyz[0].numpy(), int(m)).astype(np.int32))[None]


def _ball_query(min_radius, max_radius, nsample, xyz, center):
    """(1,N,3), (1,M,3) -> (1,M,nsample) int32, like ops/ball_query (CUDA-only there)."""
    import torch
    out = cpu.ball_query(min_radius, max_radius, int(nsample), xyz[0].numpy(), center[0].numpy())
    return torch.from_numpy(out.astype(np.int32))[None]


def fps_nn_fast_fn():
    global _fn
    if _fn is None:
        import torch
        _fn = load_def(REF_ENCODER, 'fps_NN_fast',
                       {'torch': torch, 'furthest_point_sample': _fps, 'ball_query': _ball_query})
    return _fn


def fps_nn_fast(query, key, fps_num, radius, max_cluster_samples, dist_thresh):
    """query (Q,4) / key (Nk,4) int32 (b,z,y,x) of one sample -> (Q,) int64, -1 = unassigned."""
    import torch
    q, k = torch.from_numpy(np.ascontiguousarray(query)), torch.from_numpy(np.ascontiguousarray(key))
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        return fps_nn_fast_fn()(None, q, k, fps_num, radius, max_cluster_samples, dist_thresh).numpy()
    finally:
        torch.set_num_threads(threads)
