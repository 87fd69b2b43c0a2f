"""TEST INFRASTRUCTURE for sessions without GPU time: runs the `-m gpu` tests that fit on the CPU emulation of the
CUDA kernels (MSMD_EMULATE=1, tests/conftest.py -> tests/tools/cuda_emul: every translation unit except the
thread-block-cluster kernels of points.cu, tensor-core kernels on the host model of tcgen05).

    python tests/tools/run_gpu_tests_on_emulator.py [--file test_zz_train_gpu.py] [-k EXPR] [pytest args...]

What it is good for: the Python side of the GPU tests (their own code, the wrappers, modules, autograd, plan
builder) and the kernels' logic, before a hardware run exists.  What it is not: a substitute for that run (no
timing, no memory-model races, emulated tcgen05).  Left out by default: tests that need FPS / ball query
(cluster kernels), pinned-memory / stream plumbing of the detector, and BASELINE-size scenes (minutes each on the
emulator).  Sizes in the remaining tests are what the GPU runs, so expect ~10-15 minutes for the whole file.
`-k lc_train_step` on its own (MSMD_EMULATE=1 python -m pytest tests/test_zz_train_gpu.py -m gpu -k lc_train_step) runs the
whole train step at a reduced scene size in ~16 minutes and passes.
"""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)

SKIP = ('lc_train_step or full_size or batched_dropins or executor_with_mask_sort or sparse_encoder_within_parity or fps or ball_query or nn_search or fps_nn or msmd_voxel_space or '
        'lift or depth_canvas or modality_split or native_executor_equals or config1 or hard_voxelize_full')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--file', default='test_zz_train_gpu.py')
    ap.add_argument('-k', default=None)
    args, rest = ap.parse_known_args()
    expr = 'not (%s)' % SKIP
    if args.k:
        expr = '(%s) and %s' % (args.k, expr)
    env = dict(os.environ, MSMD_EMULATE='1')
    cmd = [sys.executable, '-m', 'pytest', os.path.join(TESTS, args.file), '-q', '-m', 'gpu', '-k', expr] + rest
    print(' '.join(cmd))
    return subprocess.call(cmd, env=env)


if __name__ == '__main__':
    sys.exit(main())
