"""Summarise an ncu report (.ncu-rep) or a launch-list csv into a small text file for profiles/.
Usage: python tools/ncu_summary.py report gpurun_out/prof.ncu-rep > profiles/x.txt
       python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/y.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def report(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f'# ncu --set full --clock-control none, source: {path}')
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print(f'\n## kernel: {name[:150]}')
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:90s} {r[i]:>22s} {units[i]}')


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            ns = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = r[ki].split('(')[0][:110]
        tot[name][0] += 1
        tot[name][1] += ns
    total = sum(v[1] for v in tot.values())
    print(f'# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES), source: {path}')
    print(f'# {sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms total')
    for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'{ns / 1e6:10.3f} ms  {100 * ns / total:6.2f} %  x{n:<4d} {name}')


if __name__ == '__main__':
    {'report': report, 'launches': launches}[sys.argv[1]](sys.argv[2])
