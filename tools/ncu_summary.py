"""One line per kernel launch from an `ncu -i report.ncu-rep --page raw --csv` export (units normalised): duration, DRAM bytes and
throughput, SM / memory-pipe utilisation, occupancy, issue rate, tensor-pipe activity, long-scoreboard stalls.
usage: python tools/ncu_summary.py gpurun_out/r02_ncu_pixel_raw.csv"""
import csv
import sys


def scale(unit):
    u = unit.lower()
    return {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'us': 1.0, 'ns': 1e-3, 'ms': 1e3, 's': 1e6}.get(u, 1.0)


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def g(r, name):
        if name not in idx:
            return float('nan')
        try:
            return float(r[idx[name]].replace(',', '')) * scale(units[idx[name]])
        except ValueError:
            return float('nan')

    print('%-34s %8s %8s %7s %6s %6s %6s %5s %6s %6s %7s %5s' % ('kernel', 'us', 'dramMB', 'GB/s', 'dram%', 'sm%', 'mem%', 'occ%', 'ipc', 'tens%', 'lsb/iss', 'regs'))
    for r in data:
        name = r[idx['Kernel Name']].replace('<unnamed>::', '').replace('void ', '')[:34]
        dur = g(r, 'gpu__time_duration.sum')
        by = g(r, 'dram__bytes_read.sum') + g(r, 'dram__bytes_write.sum')
        print('%-34s %8.1f %8.1f %7.0f %6.1f %6.1f %6.1f %5.1f %6.2f %6.1f %7.2f %5s' % (
            name, dur, by / 1e6, by / dur / 1e3, g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            g(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'), g(r, 'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed'),
            g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'), g(r, 'sm__inst_executed.avg.per_cycle_active'),
            g(r, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active' in idx
            else g(r, 'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active'),
            g(r, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'), r[idx['launch__registers_per_thread']]))


if __name__ == '__main__':
    main()
