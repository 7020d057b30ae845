"""Condenses an .ncu-rep (from `ncu --set full --import-source on`) into a small text summary for profiles/.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, 'raw')
    hdr, units = raw[0], raw[1]
    print(f'# ncu summary of {rep} (ncu --set full --clock-control none)')
    for r in raw[2:]:
        print(f"\n## launch: {r[hdr.index('Kernel Name')][:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:88s} {r[i]:>16s} {units[i]}')
    src = page(rep, 'source')
    hrow = next((i for i, r in enumerate(src[:6]) if 'Source' in r and '# Samples' in r), None)      # the header row moves between ncu versions
    if hrow is not None and len(src) > hrow + 1:
        h = src[hrow]
        isrc, ismp, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
        stalls = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
        data = [r for r in src[hrow + 1:] if len(r) == len(h) and (r[ismp] or '0').isdigit()]      # a multi-launch report repeats the header row per launch
        tot = sum(int(r[ismp] or 0) for r in data) or 1
        print(f'\n## hottest SASS instructions of the last launch ({tot} warp samples, {len(data)} instructions)')
        for r in sorted(data, key=lambda r: -int(r[ismp] or 0))[:14]:
            st = sorted(((int(r[i] or 0), h[i][6:]) for i in stalls), reverse=True)[:2]
            print(f"{100 * int(r[ismp] or 0) / tot:5.1f}%  executed {r[iex]:>9s}  {r[isrc][:72]:72s} {st}")


if __name__ == '__main__':
    main()
