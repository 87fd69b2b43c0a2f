"""Builds libmsmd_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

The shared library lands at ``msmdfusion_b200/_C/libmsmd_b200.so`` so that it travels to
the GPU box with the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB = os.path.join(OUT_DIR, 'libmsmd_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
    '--expt-relaxed-constexpr',
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + \
        [os.path.join(HERE, '..', 'include', 'msmd_b200.h')]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


TRACE_LIB = os.path.join(OUT_DIR, 'libmsmd_b200_trace.so')


def build_trace(verbose=False):
    """Debug build with the per-role timeline of the tensor-core kernels compiled in (csrc/tc_trace.cuh,
    -DMSMD_TC_TRACE); used only by tools/tc_trace.py through MSMD_LIB.  The product library never carries it."""
    global LIB, OUT_DIR
    saved = (LIB, OUT_DIR)
    try:
        LIB, OUT_DIR = TRACE_LIB, os.path.join(saved[1], 'trace')
        os.makedirs(OUT_DIR, exist_ok=True)
        return build(force=not os.path.exists(TRACE_LIB) or needs_build(), verbose=verbose,
                     extra_flags=['-DMSMD_TC_TRACE'])
    finally:
        LIB, OUT_DIR = saved


PDL_LIB = os.path.join(OUT_DIR, 'libmsmd_b200_pdl.so')


def build_pdl(verbose=False):
    """Debug build with programmatic dependent launch of the tensor-core conv kernels (-DMSMD_TC_PDL: tc_launch in
    csrc/tc_common.cuh, griddepcontrol in the kernels); A/B it against the product library through MSMD_LIB."""
    global LIB, OUT_DIR
    saved = (LIB, OUT_DIR)
    try:
        LIB, OUT_DIR = PDL_LIB, os.path.join(saved[1], 'pdl')
        os.makedirs(OUT_DIR, exist_ok=True)
        return build(force=not os.path.exists(PDL_LIB) or needs_build(), verbose=verbose,
                     extra_flags=['-DMSMD_TC_PDL'])
    finally:
        LIB, OUT_DIR = saved


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, '-c', src, '-o', obj]
        if verbose:
            print(' '.join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out.decode())
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    link = [nvcc, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a',
            '-lcudart']
    if verbose:
        print(' '.join(link))
    subprocess.check_call(link)
    return LIB


if __name__ == '__main__':
    if '--pdl' in sys.argv:
        print(build_pdl(verbose=True))
        sys.exit(0)
    if '--trace' in sys.argv:
        print(build_trace(verbose=True))
        sys.exit(0)
    print(build(force='--force' in sys.argv, verbose=True,
                extra_flags=['-Xptxas', '-v'] if '--ptxas' in sys.argv else ()))
