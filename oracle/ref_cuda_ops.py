"""The reference's own CUDA ops, compiled UNMODIFIED from /root/reference into oracle/_ref/ for sm_100:

    ref_voxel_layer_cuda   mmdet3d/ops/voxel/src/{voxelization.cpp, voxelization_cpu.cpp, scatter_points_cpu.cpp,
                           voxelization_cuda.cu, scatter_points_cuda.cu}  (-DWITH_CUDA)   -> hard_voxelize (GPU)
    ref_fps_cuda           mmdet3d/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu
    ref_ball_query_cuda    mmdet3d/ops/ball_query/src/ball_query_cuda.cu
                           (the two torch-free .cu translation units only, by plain nvcc: their .cpp wrappers include
                           THC/THC.h, which torch 2.x no longer ships; the kernel launchers are called through ctypes
                           by their C++-mangled names -- `launcher(name)`)

TEST / BASELINE INFRASTRUCTURE (see oracle/README.md): only tests/ and bench.py's `cuda_baseline` leg load these.
No reference source is copied into this repository; only the built .so files land in oracle/_ref/ (git-ignored,
not gpurun-ignored, so they travel to the GPU box).  Together with oracle/_ref/ref_spconv1x.so (the reference's
vendored spconv-1.x, GPU path included) they are the reference CUDA path that CAN be built here: spconv v2.1.21,
which the hot path imports, is not vendored and not installable offline (BASELINE.md section 3 names this stand-in).
"""
import glob
import importlib.util
import os

from . import build as _build

REF_ROOT = _build.REF_ROOT
REF_DIR = _build.REF_DIR

_UNITS = {
    'ref_voxel_layer_cuda': ('mmdet3d/ops/voxel/src', ['voxelization.cpp', 'voxelization_cpu.cpp',
                                                      'scatter_points_cpu.cpp', 'voxelization_cuda.cu',
                                                      'scatter_points_cuda.cu'], ['-DWITH_CUDA']),
}
_NVCC_UNITS = {
    'ref_fps_cuda': ('mmdet3d/ops/furthest_point_sample/src/furthest_point_sample_cuda.cu',
                     'furthest_point_sampling_kernel_launcher'),
    'ref_ball_query_cuda': ('mmdet3d/ops/ball_query/src/ball_query_cuda.cu', 'ball_query_kernel_launcher'),
}
_MODS = {}


def so_path(name):
    hits = glob.glob(os.path.join(REF_DIR, name + '*.so'))
    return hits[0] if hits else None


def build(name, verbose=False):
    if so_path(name):
        return so_path(name)
    if name in _NVCC_UNITS:
        import subprocess
        src = os.path.join(REF_ROOT, _NVCC_UNITS[name][0])
        if not os.path.exists(src):
            return None
        os.makedirs(REF_DIR, exist_ok=True)
        dst = os.path.join(REF_DIR, name + '.so')
        cmd = [os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc'), '-O2', '-w', '-shared', '-Xcompiler', '-fPIC',
               '-gencode', 'arch=compute_100a,code=sm_100a', src, '-o', dst, '-lcudart']
        if verbose:
            print(' '.join(cmd))
        subprocess.check_call(cmd)
        return dst
    sub, files, defs = _UNITS[name]
    src = os.path.join(REF_ROOT, sub)
    if not os.path.isdir(src):
        return None
    from torch.utils.cpp_extension import load
    bd = os.path.join(REF_DIR, '_obj_' + name)
    os.makedirs(bd, exist_ok=True)
    os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
    load(name=name, sources=[os.path.join(src, f) for f in files], build_directory=bd,
         extra_cflags=['-O2', '-w'] + defs, extra_cuda_cflags=['-O2', '-w'] + defs, with_cuda=True, verbose=verbose)
    so = glob.glob(os.path.join(bd, name + '*.so'))
    if not so:
        return None
    dst = os.path.join(REF_DIR, os.path.basename(so[0]))
    os.replace(so[0], dst)
    return dst


def build_all(verbose=False):
    out = {}
    for name in list(_UNITS) + list(_NVCC_UNITS):
        try:
            out[name] = build(name, verbose)
        except Exception as e:  # noqa: BLE001 - an optional baseline
            out[name] = None
            print('oracle/_ref %s not built: %s' % (name, str(e)[-400:]))
    return out


def module(name):
    if name not in _MODS:
        path = so_path(name) or build(name)
        if path is None:
            return None
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _MODS[name] = mod
    return _MODS[name]


def launcher(name):
    """ctypes handle of the kernel launcher of a plain-nvcc unit (FPS / ball query), found by its mangled name."""
    import ctypes
    import subprocess
    if name in _MODS:
        return _MODS[name]
    path = so_path(name) or build(name)
    if path is None:
        return None
    want = _NVCC_UNITS[name][1]
    sym = None
    for ln in subprocess.check_output(['nm', '-D', '--defined-only', path], text=True).splitlines():
        f = ln.split()
        if len(f) == 3 and f[1] == 'T' and want in f[2] and 'with_dist' not in f[2]:
            sym = f[2]
    if sym is None:
        return None
    fn = getattr(ctypes.CDLL(path), sym)
    fn.restype = None
    _MODS[name] = fn
    return fn


if __name__ == '__main__':
    print(build_all(verbose=True))
