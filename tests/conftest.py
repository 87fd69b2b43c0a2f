import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    # the oracle is the checker for every test; building it is cheap (gcc, ~1 s)
    from oracle import build
    build.build_port()


GPU_TEST_TIMEOUT_S = int(os.environ.get('MSMD_GPU_TEST_TIMEOUT', '600'))  # per `-m gpu` test, on hardware only
EMULATE = os.environ.get('MSMD_EMULATE', '0') not in ('', '0')
# MSMD_EMULATE=1: run the `-m gpu` tests on the CPU emulation of the CUDA kernels (tests/tools/cuda_emul: every
# translation unit except the thread-block-cluster kernels of points.cu).  Test infrastructure for sessions without
# GPU time: it checks the PYTHON side of the GPU tests (wrappers, modules, autograd, the tests' own code) and the
# kernels' logic; it is not a substitute for the hardware run.  tests/tools/run_gpu_tests_on_emulator.py picks the
# tests that fit (no FPS / ball query, no full-size scenes).


class _EmuLibrary:
    """``_cabi.lib()`` look-alike over the emulated image: msmd_X -> emu_msmd_X with _cabi's own signatures."""

    def __init__(self, cabi, cdll):
        self._cabi, self._cdll, self._cache = cabi, cdll, {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            fn = getattr(self._cdll, 'emu_' + name)
            fn.restype, fn.argtypes = self._cabi.SIGNATURES[name]
            self._cache[name] = fn
        return fn


@pytest.fixture(scope='session', autouse=True)
def _emulated_library():
    if not EMULATE:
        yield None
        return
    import ctypes
    import importlib.util
    import torch
    from msmdfusion_b200 import _cabi, executor, ops
    spec = importlib.util.spec_from_file_location('emul_build', os.path.join(ROOT, 'tests', 'tools', 'cuda_emul', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cdll = ctypes.CDLL(mod.build_full())
    cdll.emu_msmd_last_error.restype = ctypes.c_char_p
    shim = _EmuLibrary(_cabi, cdll)

    def ptr(t):
        if t is None:
            return None
        assert t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())

    class Scratch:
        def get(self, device, nbytes, slot='ws'):
            return torch.empty(max(int(nbytes), 16), dtype=torch.uint8)

    saved = []
    for m_ in (_cabi, ops, executor):
        for name, val in (('lib', lambda: shim), ('ptr', ptr), ('stream', lambda device=None: None)):
            if hasattr(m_, name):
                saved.append((m_, name, getattr(m_, name)))
                setattr(m_, name, val)
    for m_ in (_cabi, ops):
        saved.append((m_, 'scratch', m_.scratch))
        m_.scratch = Scratch()
    saved.append((torch.cuda, 'synchronize', torch.cuda.synchronize))
    torch.cuda.synchronize = lambda *a, **k: None
    saved.append((_cabi, 'require_cuda', _cabi.require_cuda))
    _cabi.require_cuda = lambda t, what: None
    # the one piece that cannot be emulated (FPS runs on a thread-block cluster): the whole fps_NN_fast of a
    # sample is answered by the oracle's pinned restatement; ball query / NN search / group assignment with it
    from msmdfusion_b200 import fusion_encoder
    from oracle import cpu as _oracle

    def fps_nn_fast(query, key, fps_num, radius, max_cluster_samples, dist_thresh, base=0):
        if query.shape[0] == 0:
            return torch.empty((0,), dtype=torch.int64)
        out = _oracle.fps_nn_fast(query.cpu().numpy(), key.cpu().numpy(), fps_num, radius, max_cluster_samples,
                                  dist_thresh)
        out = torch.from_numpy(np.asarray(out, np.int64))
        return torch.where(out >= 0, out + int(base), out)
    saved.append((fusion_encoder, 'fps_nn_fast', fusion_encoder.fps_nn_fast))
    fusion_encoder.fps_nn_fast = fps_nn_fast
    yield shim
    for m_, name, val in saved:
        setattr(m_, name, val)


@pytest.fixture(autouse=True)
def _emulated_device(request, monkeypatch):
    if EMULATE and hasattr(request.module, 'dev'):
        import torch
        monkeypatch.setattr(request.module, 'dev', lambda: torch.device('cpu'))
    yield


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available() and not EMULATE and config.pluginmanager.hasplugin('timeout'):
        # a kernel that never returns must end the test process, not hold the GPU box until the caller's limit:
        # method 'thread' dumps the stacks and os._exit()s, which also tears the CUDA context down
        for item in items:
            if 'gpu' in item.keywords and item.get_closest_marker('timeout') is None:
                item.add_marker(pytest.mark.timeout(GPU_TEST_TIMEOUT_S, method='thread'))
    if torch.cuda.is_available() or EMULATE:
        return
    skip =pytest.mark.skip(reason='no CUDA device in this container (GPU tests run under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Observed absolute feature errors of the GPU parity tests (tests/test_gpu_parity.py:feat_err) -> one JSON file
    under gpurun_out/ (per test: the largest absolute error seen and the tensor's scale there)."""
    mod = sys.modules.get('test_gpu_parity')
    log = getattr(mod, 'PARITY_LOG', None) if mod is not None else None
    if not log:
        return
    import json
    worst = {}
    for test, abs_err, scale in log:
        if test not in worst or abs_err > worst[test][0]:
            worst[test] = (abs_err, scale)
    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'parity_abs_err.json'), 'w') as f:
            json.dump({'bound': 'absolute 1e-4 where max|ref| <= 10, 1e-5 of max|ref| above',
                       'overall_max_abs_err': max(v[0] for v in worst.values()),
                       'tests': {k: {'max_abs_err': v[0], 'max_abs_ref': v[1]} for k, v in sorted(worst.items())}},
                      f, indent=1)
    except OSError:
        pass
