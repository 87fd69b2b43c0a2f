"""Debug harness (GPU): tensor-core sparse conv (csrc/spconv_tc.cu) against the exact-fp32 SIMT
kernel on random rulebooks.  Prints an error summary per shape; exits non-zero on mismatch."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msmdfusion_b200 import ops  # noqa: E402


def run(cin, cout, n_in, n_out, kvol, density, seed, epilogue):
    g = torch.Generator(device='cpu').manual_seed(seed)
    dev = torch.device('cuda:0')
    feat = torch.randn(n_in, cin, generator=g).to(dev)
    w = (torch.randn(cout, kvol, cin, generator=g) / (cin * kvol * density) ** 0.5).to(dev)
    pair = torch.randint(0, n_in, (kvol, n_out), generator=g, dtype=torch.int32)
    mask = torch.rand(kvol, n_out, generator=g) < density
    pair = torch.where(mask, pair, torch.full_like(pair, -1)).to(dev)
    if kvol > 2:
        pair[1, :] = -1  # a kernel offset nobody uses -> skipped chunk(s)
    scale = shift = residual = None
    relu = False
    if epilogue:
        scale = (torch.rand(cout, generator=g) + 0.5).to(dev)
        shift = (torch.randn(cout, generator=g) * 0.1).to(dev)
        residual = torch.randn(n_out, cout, generator=g).to(dev)
        relu = True
    ref = ops.spconv_fwd(feat, ops.pack_weight(w.view(cout, kvol, 1, 1, cin)), pair, scale, shift, residual, relu)
    got = ops.spconv_fwd_tc(feat, ops.pack_weight_tc(w.view(cout, kvol, 1, 1, cin)), pair, scale, shift, residual, relu)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    mx = float(err.max())
    tag = f'cin={cin:4d} cout={cout:4d} n_in={n_in:6d} n_out={n_out:6d} kvol={kvol:2d} dens={density:.2f} epi={int(epilogue)}'
    ok = mx < 4e-5 * max(1.0, float(ref.abs().max()))  # fp32 accumulation over K <= 5184 terms: ~2e-5 relative
    print(f'{"OK  " if ok else "FAIL"} {tag}  max|err|={mx:.3e}  max|ref|={float(ref.abs().max()):.3f}')
    if not ok:
        bad = (err > 1e-4).nonzero()
        print('   bad elements:', bad.shape[0], 'of', err.numel())
        if bad.shape[0]:
            rows = torch.unique(bad[:, 0])
            cols = torch.unique(bad[:, 1])
            print('   bad rows (first 16):', rows[:16].tolist(), ' rows%128:', torch.unique(rows % 128)[:16].tolist())
            print('   bad cols (first 32):', cols[:32].tolist())
            r, c = int(bad[0, 0]), int(bad[0, 1])
            print(f'   e.g. [{r},{c}] got {float(got[r, c]):.6f} ref {float(ref[r, c]):.6f}')
    return ok


def main():
    shapes = [
        # cin, cout, n_in, n_out, kvol, density, epilogue
        (32, 32, 1000, 128, 1, 1.0, False),     # one chunk, one tile: the minimal case
        (32, 32, 1000, 128, 3, 1.0, False),
        (32, 32, 5000, 1000, 27, 0.3, False),
        (16, 16, 20000, 19000, 27, 0.1, True),
        (5, 16, 20000, 19000, 27, 0.1, True),   # scalar-gather path (cin % 4 != 0)
        (64, 64, 30000, 30000, 27, 0.4, True),
        (128, 128, 30000, 25000, 27, 0.5, True),
        (80, 80, 30000, 30011, 27, 0.3, True),
        (96, 128, 30000, 7000, 27, 0.5, True),
        (192, 192, 20000, 20000, 27, 0.5, True),
        (128, 128, 20000, 9000, 3, 0.7, True),  # conv_out k=(3,1,1)
        (16, 32, 20000, 6000, 27, 0.2, True),
        (20, 24, 3000, 777, 27, 0.3, True),     # cout not a multiple of 16, odd sizes
        (32, 32, 100, 1, 27, 0.5, False),
        (128, 128, 30000, 21509, 27, 0.5, True),  # 169 tiles: split-K pairs
        (192, 192, 9000, 6000, 27, 0.5, True),    # 47 tiles: split-K pairs
        (128, 128, 9000, 130, 3, 0.7, True),      # 2 tiles, 12 chunks
    ]
    ok = True
    variants = [int(v) for v in os.environ.get('TC_VARIANTS', '3,2').split(',')]
    for variant in variants:
        ops.set_tc_variant(variant)
        print('--- variant', variant)
        for i, s in enumerate(shapes):
            try:
                ok &= run(*s[:6], seed=i, epilogue=s[6])
            except Exception as e:  # noqa: BLE001
                print('EXC ', s, repr(e))
                ok = False
                break
    ops.set_tc_variant(0)
    print('ALL OK' if ok else 'MISMATCH')
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
