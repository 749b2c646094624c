#!/usr/bin/env python
"""Condense an `ncu --page raw --csv` dump to the handful of counters DESIGN.md / bench.py cite, one block per launch."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_uniform.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('== %s  grid %s block %s' % (d.get('Kernel Name', '?')[:60], d.get('Grid Size'), d.get('Block Size')))
        for k in KEYS:
            for h in hdr:
                if h == k or h.endswith('.' + k):
                    print('   %-80s %14s %s' % (k, d[h], u[h]))
                    break


if __name__ == '__main__':
    main(sys.argv[1])
