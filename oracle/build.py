"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE, never on the product path).

* ``build_port()``  -- gcc-compiles ``oracle/c/msmd_oracle.c`` (the restatement) into
  ``oracle/_build/libmsmd_oracle.so``.
* ``build_ref()``   -- when ``/root/reference`` is present, compiles the reference's OWN
  CPU ``hard_voxelize`` from the sources where they lie
  (``mmdet3d/ops/voxel/src/{voxelization.cpp,voxelization_cpu.cpp,scatter_points_cpu.cpp}``)
  into ``oracle/_ref/``.  No reference source is copied into this repo; only the built
  ``.so`` lands in ``oracle/_ref/`` (git-ignored, but it travels to the GPU box).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
PORT_SO = os.path.join(HERE, '_build', 'libmsmd_oracle.so')
REF_DIR = os.path.join(HERE, '_ref')
REF_NAME = 'ref_voxel_layer'


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_port(verbose=False):
    src = os.path.join(HERE, 'c', 'msmd_oracle.c')
    os.makedirs(os.path.dirname(PORT_SO), exist_ok=True)
    if _newer(PORT_SO, [src]):
        return PORT_SO
    cmd = ['gcc', '-O2', '-fopenmp', '-shared', '-fPIC', '-std=c11', '-o', PORT_SO, src, '-lm']
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return PORT_SO


def ref_so_path():
    hits = glob.glob(os.path.join(REF_DIR, REF_NAME + '*.so'))
    return hits[0] if hits else None


def build_ref(verbose=False):
    """Compile the reference CPU voxelization (needs /root/reference + torch headers)."""
    if ref_so_path():
        return ref_so_path()
    src_dir = os.path.join(REF_ROOT, 'mmdet3d', 'ops', 'voxel', 'src')
    if not os.path.isdir(src_dir):
        return None
    from torch.utils.cpp_extension import load
    os.makedirs(REF_DIR, exist_ok=True)
    build_dir = os.path.join(REF_DIR, '_obj')
    os.makedirs(build_dir, exist_ok=True)
    load(
        name=REF_NAME,
        sources=[os.path.join(src_dir, f)
                 for f in ('voxelization.cpp', 'voxelization_cpu.cpp', 'scatter_points_cpu.cpp')],
        build_directory=build_dir,
        extra_cflags=['-O2'],
        verbose=verbose,
        is_python_module=True,
    )
    so = glob.glob(os.path.join(build_dir, REF_NAME + '*.so'))
    if not so:
        return None
    dst = os.path.join(REF_DIR, os.path.basename(so[0]))
    os.replace(so[0], dst)
    return dst


if __name__ == '__main__':
    print(build_port(verbose=True))
    if '--ref' in sys.argv:
        print(build_ref(verbose=True))
