"""Turn ncu output brought back in gpurun_out/ into the small tracked summaries in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_hot.csv profiles/r1_launches_hot.csv
    python profiles/summarize.py full gpurun_out/prof_hot.ncu-rep profiles/r1_ncu_full_hot.csv
    python profiles/summarize.py traffic gpurun_out/r2_prof_hot.ncu-rep profiles/r2_ncu_traffic.json "<source note>"
"""
import json
import collections
import csv
import subprocess
import sys

FULL_COLS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
             'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
             'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
             'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
             'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
             'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
             'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
             'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
             'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
             'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
             'l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed']


def _us(value, unit):
    v = float(value.replace(',', ''))
    return v * {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 's': 1e6, 'second': 1e6, 'nsecond': 1e-3}.get(unit, 1.0)


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith('=='))]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += _us(r[vi], r[ui])
    total = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel', 'launches', 'total_us', 'avg_us', 'share_pct'])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k[:110], n, '%.2f' % t, '%.2f' % (t / n), '%.2f' % (100 * t / total)])
        w.writerow(['TOTAL', sum(a[0] for a in agg.values()), '%.2f' % total, '', '100'])
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in FULL_COLS if c in hdr]
    with open(dst, 'w') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i][:100] for i in idx])
    print(open(dst).read())


def traffic(src, dst, note=""):
    """Per-launch DRAM bytes of every captured kernel (first capture of each name) -> the table bench.py reads."""
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    out = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']].split('(')[0].split('<')[0].split('::')[-1].replace('void ', '').strip()
        if name in out:
            continue
        rd = float(r[col['dram__bytes_read.sum']].replace(',', '')) * scale.get(units[col['dram__bytes_read.sum']], 1.0)
        wr = float(r[col['dram__bytes_write.sum']].replace(',', '')) * scale.get(units[col['dram__bytes_write.sum']], 1.0)
        us = _us(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
        out[name] = {'dram_bytes': int(rd + wr), 'dram_read_MB': round(rd / 1e6, 2), 'dram_write_MB': round(wr / 1e6, 2), 'ncu_time_us': round(us, 2)}
    with open(dst, 'w') as f:
        json.dump({'source': note, 'per_launch': out}, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    {'launches': launches, 'full': full, 'traffic': traffic}[sys.argv[1]](*sys.argv[2:])
