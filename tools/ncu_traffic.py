"""Turns an `ncu --set full` capture of the conv launches of ONE scene into the small JSON bench.py reads for
`roofline.traffic` (profiles/traffic_<workload>_<profile>_<precision>.json) plus a per-launch table.

    ncu --set full --clock-control none --import-source on -k regex:spconv_fwd -s <launches of pass 1> -c <launches of pass 2> \
        -o gpurun_out/prof python tools/prof_conv.py [--lc] --precision bf16x3c          (on the GPU box)
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep LC S bf16x3c                       (here: ncu reads reports without a GPU)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, workload, profile, precision = sys.argv[1:5]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}

    def val(r, name):
        x = float(r[col[name]].replace(',', ''))
        u = units[col[name]]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'usecond': 1.0, 'msecond': 1e3, 'nsecond': 1e-3,
                 'second': 1e6, '%': 1.0}.get(u, 1.0)
        return x * scale

    launches = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        launches.append(dict(kernel=r[col['Kernel Name']][:48], grid=r[col['launch__grid_size']],
                             us=val(r, 'gpu__time_duration.sum'),
                             dram_read=val(r, 'dram__bytes_read.sum'), dram_write=val(r, 'dram__bytes_write.sum'),
                             tensor_active_pct=val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                             warps_active_pct=val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'),
                             regs=int(float(r[col['launch__registers_per_thread']]))))
    tot_us = sum(l['us'] for l in launches)
    doc = dict(source='dram__bytes_read.sum + dram__bytes_write.sum summed over the %d conv launches of one scene, '
                      '`ncu --set full --clock-control none` capture %s (cold-cache, serialised launches)' % (
                          len(launches), os.path.basename(rep)),
               workload=workload, profile=profile, precision=precision, launches=len(launches),
               dram_bytes_per_step=sum(l['dram_read'] + l['dram_write'] for l in launches),
               sum_time_us=tot_us,
               tensor_active_pct_time_weighted=round(sum(l['tensor_active_pct'] * l['us'] for l in launches) / max(tot_us, 1e-9), 2),
               per_launch=launches)
    dst = os.path.join(ROOT, 'profiles', 'traffic_%s_%s_%s.json' % (workload, profile, precision))
    json.dump(doc, open(dst, 'w'), indent=1)
    print(dst, 'launches', len(launches), 'dram MB', round(doc['dram_bytes_per_step'] / 1e6, 1), 'sum us', round(tot_us, 1),
          'tensor%', doc['tensor_active_pct_time_weighted'])


if __name__ == '__main__':
    main()
