"""Per-layer roofline table from a bench.py --breakdown file (algorithmic bytes / flops per launch,
SURVEY.md section 8(d), against the measured peaks in MEASURED_PEAKS.json)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, out):
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}
    hbm, tf32 = pk['hbm_gbs'], pk['bf16_tflops'] / 2.0
    b = json.load(open(path))
    rows = [r for r in b['first_step_calls'] if r['op'] == 'spconv_fwd']
    lines = ['| layer | N_out | pairs | ms | algorithmic MB | GB/s | % HBM peak | algorithmic TFLOP/s | % tf32 peak '
             '(x3 MMAs issued) | binding roof (3xTF32) |', '|---|---|---|---|---|---|---|---|---|---|']
    for r in rows:
        by = 4.0 * (r['n_in'] * r['cin'] + r['n_out'] * r['cout'] + r['kvol'] * r['cin'] * r['cout']) + \
            4.0 * r['kvol'] * r['n_out'] + (4.0 * r['n_out'] * r['cout'] if r['residual'] else 0.0)
        fl = 2.0 * r['pairs'] * r['cin'] * r['cout']
        t = r['ms'] * 1e-3
        gbs, tfl = by / t / 1e9, fl / t / 1e12
        t_hbm, t_tc = by / (hbm * 1e9), 3.0 * fl / (tf32 * 1e12)
        roof = 'tensor' if t_tc > t_hbm else 'HBM'
        frac = max(t_hbm, t_tc) / t
        lines.append('| %d→%d k%d%s | %d | %d | %.4f | %.1f | %.0f | %.1f | %.1f | %.1f (%.1f) | %s: %.0f %% of roof |' % (
            r['cin'], r['cout'], r['kvol'], '+res' if r['residual'] else '', r['n_out'], r['pairs'], r['ms'], by / 1e6,
            gbs, 100 * gbs / hbm, tfl, 100 * tfl / tf32, 100 * 3 * tfl / tf32, roof, 100 * frac))
    open(out, 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
