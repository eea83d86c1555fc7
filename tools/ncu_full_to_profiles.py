#!/usr/bin/env python
"""From `ncu -i <rep> --page raw --csv` of ONE step (one launch per kernel of b200mrc_decompose) write
  <out>.csv           the per-kernel summary table kept under profiles/ (duration, grid, registers, smem, warps active,
                      issue-slot and pipe utilisation, executed instructions, DRAM bytes), and
  profiles/traffic.json   dram__bytes_read.sum + dram__bytes_write.sum per kernel and for the step (bench.py copies these
                      into roofline.per_kernel[*].dram_bytes / roofline.traffic).

  python tools/ncu_full_to_profiles.py raw.csv profiles/r2u_ncu_full_summary_64pages.csv [--traffic profiles/traffic.json]
"""
import csv, json, re, sys

COLS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
UNIT_BYTES = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main():
    raw, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[sys.argv.index('--traffic') + 1] if '--traffic' in sys.argv else None
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]].replace('b200mrc::<', '') if c == 'Kernel Name' else r[idx[c]] for c in cols])
    if traffic_path:
        tr, total = {}, 0.0
        for r in rows[2:]:
            name = re.search(r'k_[a-z0-9_]+', r[idx['Kernel Name']]).group(0)
            b = sum(float(r[idx[c]]) * UNIT_BYTES[units[idx[c]]] for c in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            tr[name] = tr.get(name, 0.0) + b
            total += b
        tr = {k: round(v, -3) for k, v in tr.items()}
        tr['step_total'] = round(total, -3)
        tr['source'] = ('%s (ncu --set full --clock-control none, one b200mrc_decompose step on 64 pages 3300x2550 RGB, '
                        'dram__bytes_read.sum + dram__bytes_write.sum per kernel; step_total = all kernels of one step)' % out)
        json.dump(tr, open(traffic_path, 'w'), indent=1)
        print(json.dumps(tr))


if __name__ == '__main__':
    main()
