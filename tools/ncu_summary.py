"""Summarise ncu outputs into small text files for profiles/:
    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active']


def launches(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r['Kernel Name'].split('(')[0].replace('void ', '').replace('<unnamed>::', '')
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value']) / 1e6
    tot = sum(v[1] for v in agg.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)')
    print('%-44s %5s %11s %7s %11s' % ('kernel', 'n', 'total ms', 'share', 'ms/launch'))
    for k, v in agg.items():
        print('%-44s %5d %11.3f %6.1f%% %11.4f' % (k[:44], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    print('%-44s %5d %11.3f' % ('total', sum(v[0] for v in agg.values()), tot))


def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print('# ncu --set full --clock-control none; source: %s' % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(d.get('Kernel Name', '?')[:150])
        for k in KEYS:
            if k in d:
                print('    %-82s %s %s' % (k, d[k], u[k]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'report': report}[sys.argv[1]](sys.argv[2])
