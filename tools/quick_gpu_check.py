"""GPU smoke check WITHOUT torch (ctypes + libcudart + numpy only, a few seconds of GPU time): runs the kernels
that were written without hardware access through the C ABI on small problems and compares them with numpy
restatements.  Meant as the first command of a GPU call (`import torch` alone costs a fresh box up to a minute):

    python tools/quick_gpu_check.py [--out gpurun_out/quick_check.json]

Checks: 3xTF32 forward (the GPU-verified path, as a sanity anchor), bf16x3 / bf16 forward (csrc/spconv_tc16.cu),
mask-sorted forward (msmd_rulebook_mask_sort + msmd_spconv_fwd_tc_sorted), data gradient through the forward
kernels, SIMT and tensor-core weight gradient.  Exit code 0 = all inside tolerance.  Not a test of the suite
(tests/ are); a fast first look.
"""
import argparse
import ctypes
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cudart():
    cands = ['/usr/local/cuda/lib64/libcudart.so'] + sorted(glob.glob('/usr/local/cuda/lib64/libcudart.so.*')) + \
        sorted(glob.glob(os.path.join(sys.prefix, 'lib/python*/site-packages/nvidia/cuda_runtime/lib/libcudart.so*')))
    for c in cands:
        if os.path.exists(c):
            return ctypes.CDLL(c, mode=ctypes.RTLD_GLOBAL)
    raise RuntimeError('libcudart not found')


class Dev:
    """Minimal device-memory helper over the CUDA runtime."""

    def __init__(self):
        self.rt = load_cudart()
        self.rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
        self.rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.rt.cudaMemset.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t]
        self.rt.cudaGetErrorString.restype = ctypes.c_char_p
        self.bufs = []

    def ok(self, st, what):
        if st != 0:
            raise RuntimeError('%s: %s' % (what, self.rt.cudaGetErrorString(st).decode()))

    def alloc(self, nbytes, fill=None):
        p = ctypes.c_void_p()
        self.ok(self.rt.cudaMalloc(ctypes.byref(p), max(int(nbytes), 256)), 'cudaMalloc')
        if fill is not None:
            self.ok(self.rt.cudaMemset(p, fill, max(int(nbytes), 256)), 'cudaMemset')
        self.bufs.append(p)
        return p

    def put(self, a):
        a = np.ascontiguousarray(a)
        p = self.alloc(a.nbytes)
        self.ok(self.rt.cudaMemcpy(p, a.ctypes.data_as(ctypes.c_void_p), a.nbytes, 1), 'H2D')
        return p

    def get(self, p, shape, dtype):
        out = np.empty(shape, dtype)
        self.ok(self.rt.cudaDeviceSynchronize(), 'sync (a kernel failed)')
        self.ok(self.rt.cudaMemcpy(out.ctypes.data_as(ctypes.c_void_p), p, out.nbytes, 2), 'D2H')
        return out


def subm_pairs(idx, shape):
    """pair_fwd (27, n) of a 3x3x3 SubM convolution: k = (kz*3+ky)*3+kx reads the voxel at o + (k - centre)."""
    D, H, W = shape
    lin = (idx[:, 0] * D + idx[:, 1]) * H * W + idx[:, 2] * W + idx[:, 3]
    table = {int(v): i for i, v in enumerate(lin)}
    n = idx.shape[0]
    pair = np.full((27, n), -1, np.int32)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                z, y, x = idx[:, 1] + dz, idx[:, 2] + dy, idx[:, 3] + dx
                okm = (z >= 0) & (z < D) & (y >= 0) & (y < H) & (x >= 0) & (x < W)
                l2 = (idx[:, 0] * D + z) * H * W + y * W + x
                pair[k] = [table.get(int(v), -1) if o else -1 for v, o in zip(l2, okm)]
                k += 1
    return pair


def conv_ref(feat, w, pair):
    """float64 restatement: out[o] = sum_k W[:, k, :] @ feat[pair[k, o]]."""
    cout, cin = w.shape[0], w.shape[-1]
    wk = w.reshape(cout, -1, cin).astype(np.float64)
    out = np.zeros((pair.shape[1], cout))
    f = feat.astype(np.float64)
    for k in range(pair.shape[0]):
        m = pair[k] >= 0
        out[m] += f[pair[k][m]] @ wk[:, k, :].T
    return out


def wgrad_ref(feat, go, pair, cout, cin):
    gw = np.zeros((cout, pair.shape[0], cin))
    f, g = feat.astype(np.float64), go.astype(np.float64)
    for k in range(pair.shape[0]):
        m = pair[k] >= 0
        gw[:, k, :] = g[m].T @ f[pair[k][m]]
    return gw


def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(1.0, np.abs(b).max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=None)
    ap.add_argument('--lib', default=os.environ.get('MSMD_LIB', os.path.join(ROOT, 'msmdfusion_b200', '_C', 'libmsmd_b200.so')))
    args = ap.parse_args()
    t0 = time.time()
    L = ctypes.CDLL(args.lib)
    L.msmd_last_error.restype = ctypes.c_char_p
    for name in ('msmd_spconv_tc_packed_floats', 'msmd_spconv_tc16_packed_bytes', 'msmd_spconv_tc_workspace',
                 'msmd_rulebook_mask_sort_workspace', 'msmd_spconv_bwd_weight_workspace',
                 'msmd_spconv_bwd_weight_tc_workspace'):
        getattr(L, name).restype = ctypes.c_size_t
    d = Dev()
    results, failed, info = [], [], []
    VP = ctypes.c_void_p

    def call(name, *a):
        fn = getattr(L, name)
        st = fn(*a)
        if st != 0:
            raise RuntimeError('%s -> %d: %s' % (name, st, L.msmd_last_error().decode()))

    def record(name, err, tol):
        results.append(dict(check=name, err=err, tol=tol, ok=bool(err < tol)))
        print('%-46s err %.3e  tol %.0e  %s' % (name, err, tol, 'ok' if err < tol else 'FAIL'), flush=True)
        if not err < tol:
            failed.append(name)

    rng = np.random.default_rng(0)
    shape = [9, 40, 40]
    for cin, cout, n in ((16, 16, 3000), (64, 64, 3000), (128, 128, 2500), (20, 144, 2000)):
        try:
            lin = rng.choice(shape[0] * shape[1] * shape[2], size=n, replace=False)
            idx = np.stack([lin * 0, lin // 1600, (lin // 40) % 40, lin % 40], 1).astype(np.int32)
            pair = subm_pairs(idx, shape)
            feat = rng.standard_normal((n, cin)).astype(np.float32)
            w = (rng.standard_normal((cout, 27, cin)) / np.sqrt(cin * 27 * 0.2)).astype(np.float32)
            go = rng.standard_normal((n, cout)).astype(np.float32)
            scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
            shift = rng.standard_normal(cout).astype(np.float32)
            res = rng.standard_normal((n, cout)).astype(np.float32)
            ref = conv_ref(feat, w, pair)
            ref_epi = np.maximum(ref * scale + shift + res, 0)
            ref16 = conv_ref(bf16_round(feat), bf16_round(w), pair)
            tag = '%d->%d n=%d' % (cin, cout, n)
            dfeat, dw, dpair, dgo = d.put(feat), d.put(w), d.put(pair), d.put(go)
            dscale, dshift, dres = d.put(scale), d.put(shift), d.put(res)
            dout = d.alloc(n * cout * 4, fill=0xFF)
            # --- 3xTF32 (anchor) ---
            ptc = d.alloc(L.msmd_spconv_tc_packed_floats(cout, 27, cin) * 4)
            call('msmd_spconv_tc_pack_weight', dw, cout, 27, cin, ptc, None)
            wsb = L.msmd_spconv_tc_workspace(n, cout)
            ws = d.alloc(wsb) if wsb else None
            call('msmd_spconv_fwd_tc_ws', dfeat, n, ptc, dpair, n, cin, cout, 27, dscale, dshift, dres, 1, dout, ws,
                 ctypes.c_size_t(wsb), None)
            plain = d.get(dout, (n, cout), np.float32)
            record('tf32x3 fwd+epilogue  ' + tag, rel(plain, ref_epi), 1e-5)
            # --- mask-sorted ---
            mb = L.msmd_rulebook_mask_sort_workspace(n)
            dperm, dps, dmw = d.alloc(n * 4), d.alloc(27 * n * 4), d.alloc(mb)
            call('msmd_rulebook_mask_sort', dpair, 27, n, dperm, dps, dmw, ctypes.c_size_t(mb), None)
            perm = d.get(dperm, (n,), np.int32)
            ps = d.get(dps, (27, n), np.int32)
            okp = np.array_equal(np.sort(perm), np.arange(n)) and np.array_equal(ps, pair[:, perm])
            record('mask sort permutation ' + tag, 0.0 if okp else 1.0, 0.5)
            d.rt.cudaMemset(dout, 0xFF, n * cout * 4)
            call('msmd_spconv_fwd_tc_sorted', dfeat, n, ptc, dps, dperm, n, cin, cout, 27, dscale, dshift, dres, 1,
                 dout, ws, ctypes.c_size_t(wsb), None)
            record('tf32x3 mask-sorted    ' + tag, rel(d.get(dout, (n, cout), np.float32), ref_epi), 1e-5)
            # --- 16-bit operand modes ---
            for x3, name, r, tol in ((1, 'bf16x3', ref, 2e-5), (0, 'bf16  ', ref16, 1e-5)):
                p16 = d.alloc(L.msmd_spconv_tc16_packed_bytes(cout, 27, cin, x3))
                call('msmd_spconv_tc16_pack_weight', dw, cout, 27, cin, x3, p16, None)
                d.rt.cudaMemset(dout, 0xFF, n * cout * 4)
                call('msmd_spconv_fwd_tc16', dfeat, n, p16, dpair, None, n, cin, cout, 27, x3, None, None, None, 0,
                     dout, None)
                record('%s fwd            %s' % (name, tag), rel(d.get(dout, (n, cout), np.float32), r), tol)
                d.rt.cudaMemset(dout, 0xFF, n * cout * 4)
                call('msmd_spconv_fwd_tc16', dfeat, n, p16, dps, dperm, n, cin, cout, 27, x3, dscale, dshift, dres, 1,
                     dout, None)
                r_epi = np.maximum(r * scale + shift + res, 0)
                record('%s sorted+epilogue %s' % (name, tag), rel(d.get(dout, (n, cout), np.float32), r_epi), tol)
            # --- variant 3 of the 16-bit modes (A operand in tensor memory): informational, layout unconfirmed ---
            try:
                L.msmd_spconv_tc16_workspace.restype = ctypes.c_size_t
                call('msmd_spconv_tc16_set_variant', 3)
                for x3, name, r, tol in ((1, 'bf16x3', ref, 2e-5), (0, 'bf16  ', ref16, 1e-5)):
                    p16 = d.alloc(L.msmd_spconv_tc16_packed_bytes(cout, 27, cin, x3))
                    call('msmd_spconv_tc16_pack_weight', dw, cout, 27, cin, x3, p16, None)
                    wb3 = L.msmd_spconv_tc16_workspace(n, cout)
                    ws3 = d.alloc(wb3, fill=0) if wb3 else None
                    d.rt.cudaMemset(dout, 0xFF, n * cout * 4)
                    call('msmd_spconv_fwd_tc16_ws', dfeat, n, p16, dpair, None, n, cin, cout, 27, x3, None, None, None,
                         0, dout, ws3, ctypes.c_size_t(wb3), None)
                    e3 = rel(d.get(dout, (n, cout), np.float32), r)
                    info.append(dict(check='%s variant 3 (TMEM A) %s' % (name, tag), err=e3, tol=tol, ok=bool(e3 < tol)))
                    print('%-46s err %.3e  tol %.0e  %s (informational)' % (info[-1]['check'], e3, tol,
                                                                            'ok' if e3 < tol else 'MISMATCH'), flush=True)
            finally:
                call('msmd_spconv_tc16_set_variant', 2)
            # --- weight gradient: SIMT (exact fp32) and tensor-core ---
            gref = wgrad_ref(feat, go, pair, cout, cin)
            dgw = d.alloc(cout * 27 * cin * 4, fill=0xFF)
            for tc, name in ((0, 'wgrad simt'), (1, 'wgrad tc  ')):
                if cin % 4 or cout % 4:
                    continue
                call('msmd_spconv_set_wgrad_tc', tc)
                wb = L.msmd_spconv_bwd_weight_workspace(n, cin, cout, 27)
                dws = d.alloc(wb)
                call('msmd_spconv_bwd_weight', dfeat, n, dgo, dpair, n, cin, cout, 27, dgw, dws, ctypes.c_size_t(wb), None)
                record('%s           %s' % (name, tag), rel(d.get(dgw, (cout, 27, cin), np.float32), gref), 1e-4)
            call('msmd_spconv_set_wgrad_tc', 0)
            # --- data gradient (SubM: pair_fwd with the transposed, k-reversed weight) ---
            dwt = d.alloc(cout * 27 * cin * 4)
            call('msmd_spconv_transpose_weight', dw, cout, 27, cin, 1, dwt, None)
            pt = d.alloc(L.msmd_spconv_tc_packed_floats(cin, 27, cout) * 4)
            call('msmd_spconv_tc_pack_weight', dwt, cin, 27, cout, pt, None)
            dgi = d.alloc(n * cin * 4, fill=0xFF)
            wsb2 = L.msmd_spconv_tc_workspace(n, cin)
            ws2 = d.alloc(wsb2) if wsb2 else None
            call('msmd_spconv_bwd_data', dgo, n, pt, 1, dpair, n, cin, cout, 27, dgi, ws2, ctypes.c_size_t(wsb2), None)
            wt = np.ascontiguousarray(w.reshape(cout, 27, cin)[:, ::-1, :].transpose(2, 1, 0))   # [ci][k'][co]
            record('dgrad (tc, SubM mirror) ' + tag, rel(d.get(dgi, (n, cin), np.float32), conv_ref(go, wt, pair)), 1e-5)
        except Exception as e:   # keep going: every shape is an independent look
            print('EXCEPTION %d->%d: %s' % (cin, cout, e), flush=True)
            failed.append('%d->%d: %s' % (cin, cout, e))
    summary = dict(seconds=round(time.time() - t0, 2), failed=failed, results=results, informational=info)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(summary, open(args.out, 'w'), indent=1)
    print('quick_gpu_check: %d checks, %d failed, %.1f s' % (len(results), len(failed), time.time() - t0))
    return 1 if failed else 0


if __name__ == '__main__':
    sys.exit(main())
