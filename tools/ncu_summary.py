"""Summarise ncu outputs into small CSV / markdown files for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches_r2.csv profiles/r01_launch_shares_bench_r2.md
    python tools/ncu_summary.py full gpurun_out/prof_bench_r2.ncu-rep profiles/r01_ncu_full_bench_r2.csv
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes.sum.per_second',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
]


def short(name):
    name = re.sub(r'void |sfem::|\(anonymous namespace\)::|<?unnamed>::', '', name)
    m = re.match(r'([A-Za-z_0-9:]+(<[^(]*>)?)', name)
    return (m.group(1) if m else name)[:70]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith('==')]
    hdr = rows[0]
    ik, iv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    iu = hdr.index('Metric Unit')
    agg = defaultdict(lambda: [0, 0.0])
    tot = 0.0
    body = rows[1:]
    # bench.py brackets its timed region with two erfinv marker kernels: keep what lies between them, and only
    # this library's kernels (the L2 flush between steps is torch)
    marks = [i for i, r in enumerate(body) if 'erfinv' in r[ik]]
    note = ''
    if len(marks) >= 2:
        body = [r for r in body[marks[0] + 1:marks[1]] if 'sfem' in r[ik]]
        note = ' (timed region between the two marker kernels, libsulcusfem kernels only)'
    elif len(marks) == 1:
        body = [r for r in body[marks[0] + 1:] if 'sfem' in r[ik]]
        note = ' (timed region from the first marker kernel to the end of the capture -- capture cut short, libsulcusfem kernels only)'
    for r in body:
        try:
            v = float(r[iv].replace(',', ''))
        except ValueError:
            continue
        if r[iu] == 'ns':
            v /= 1e3
        elif r[iu] == 'ms':
            v *= 1e3
        k = short(r[ik])
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    with open(dst, 'w') as f:
        f.write(f"# ncu launch list summary ({src}){note}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms kernel time\n\n")
        f.write("ncu times are cold-cache and serialised (one kernel at a time): compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.2f} |\n")
    print("wrote", dst)


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    ik = hdr.index('Kernel Name')
    with open(dst, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + [f"{k} [{units[i]}]" for k, i in cols])
        for r in rows[2:]:
            w.writerow([short(r[ik])] + [r[i] for _, i in cols])
    print("wrote", dst, len(rows) - 2, "kernels")


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
