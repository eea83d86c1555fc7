#!/usr/bin/env python
"""Summary of one `ncu --set full --import-source on` capture of one kernel: headline metrics, stall reasons, and the
source lines that hold most samples / most executed instructions.

    ncu -i cap.ncu-rep --page raw --csv > raw.csv
    ncu -i cap.ncu-rep --page source --csv --print-source cuda,sass > src.csv     (or --print-source cuda)
    python tools/ncu_source_summary.py raw.csv src.csv <rows processed by one launch, e.g. pages*strips*H warp-rows>
"""
import csv,re,collections,sys
raw=sys.argv[1]; cs=sys.argv[2]; nrows=float(sys.argv[3])
rows=list(csv.reader(open(raw)))
hdr=rows[0]; r=rows[2]
idx={h:i for i,h in enumerate(hdr)}
keys=['gpu__time_duration.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__warps_eligible.avg.per_cycle_active','launch__grid_size','launch__block_size','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in idx: print('%-75s %s'%(k,r[idx[k]]))
rows=list(csv.reader(open(cs)))
cur_file=None
agg={}; stall_tot=collections.Counter()
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; sidx=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if r[0]!='' and r[0].isdigit():
        try: samples=int(r[4]); inst=int(r[7])
        except: continue
        st=collections.Counter()
        for i,h in sidx:
            try: st[h]+=int(r[i])
            except: pass
        key=(cur_file,int(r[0]))
        if key not in agg: agg[key]=[r[1],0,0,collections.Counter()]
        agg[key][1]+=samples; agg[key][2]+=inst; agg[key][3].update(st)
        stall_tot.update(st)
tot_s=sum(v[1] for v in agg.values()); tot_i=sum(v[2] for v in agg.values())
print('stalls',{k[6:]:round(100*v/tot_s,1) for k,v in stall_tot.most_common(10)})
print('inst per warp-row: %.0f'%(tot_i/nrows))
print('--- top by samples')
for (f,l),(src,s,i,st) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    top=', '.join('%s %.1f'%(k[6:],100*v/tot_s) for k,v in st.most_common(2))
    print('%-16s %4d samp %5.1f%% inst %5.1f%% [%s]  %s'%(f[:16],l,100*s/tot_s,100*i/tot_i,top,src.strip()[:70]))
print('--- top by inst')
for (f,l),(src,s,i,st) in sorted(agg.items(), key=lambda kv:-kv[1][2])[:28]:
    print('%-16s %4d inst %5.1f%% (%5.1f/warp-row) samp %5.1f%%  %s'%(f[:16],l,100*i/tot_i,i/nrows,100*s/tot_s,src.strip()[:80]))
