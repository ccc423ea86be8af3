#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: python profiles/summarize.py launches <launches.csv>            -> per-kernel share of device time
       python profiles/summarize.py report <file.ncu-rep> [kernel re]  -> key metrics of each captured launch"""
import collections, csv, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_wait.ratio', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct']


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[-70:]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print('# per-kernel device time from `ncu --metrics gpu__time_duration.sum` (cold-cache, serialised: compare shares)')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-72s n=%4d  %11.1f us  %5.1f%%' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))


def report(path, pattern=None):
    cmd = ['ncu', '-i', path, '--page', 'raw', '--csv']
    if pattern:
        cmd += ['--kernel-name', 'regex:' + pattern]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('## %s' % r[hdr.index('Kernel Name')][:110])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('%-78s %18s %s' % (k, r[i], units[i]))
        print()


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
