#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            print(w, units[hdr.index(w)], r[hdr.index(w)][:100])
    st = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for h, v in st if v > 0.05]
    st.sort(key=lambda x: -x[1])
    print('stalls per issue:', ', '.join('%s=%.2f' % x for x in st[:9]))
    pipes = [(h.replace('sm__inst_executed_pipe_', '').replace('.avg.pct_of_peak_sustained_active', ''), float(r[i])) for i, h in enumerate(hdr) if h.startswith('sm__inst_executed_pipe_') and h.endswith('.avg.pct_of_peak_sustained_active')]
    pipes.sort(key=lambda x: -x[1])
    print('pipes %:', ', '.join('%s=%.1f' % x for x in pipes[:8]))
