#!/bin/bash
# usage: benchmarks/ncu_summary.sh <file.ncu-rep>   -- headline metrics, dynamic instruction mix and stall reasons
f=$1
ncu -i $f --page details --csv 2>/dev/null | python -c "
import csv,sys
r=csv.reader(sys.stdin)
hdr=next(r)
iS=hdr.index('Section Name'); iM=hdr.index('Metric Name'); iU=hdr.index('Metric Unit'); iV=hdr.index('Metric Value'); iK=hdr.index('Kernel Name')
keep=('Duration','Memory Throughput','DRAM Throughput','L1/TEX Cache Throughput','L2 Cache Throughput','Compute (SM) Throughput','Issue Slots Busy','Executed Ipc Active','Registers Per Thread','Achieved Occupancy','Theoretical Occupancy','Eligible Warps Per Scheduler','No Eligible','Block Limit Shared Mem','Block Limit Registers','Mem Pipes Busy','Max Bandwidth','Mem Busy','Grid Size','Block Size','Dynamic Shared Memory Per Block','SM Frequency')
first=True
for row in r:
    if first: print('kernel:',row[iK][:120]); first=False
    if row[iM] in keep: print(' ',row[iS],'|',row[iM],'|',row[iU],'|',row[iV])
"
ncu -i $f --page source --csv 2>/dev/null > /tmp/_src.csv; python $(dirname $0)/ncu_mix.py /tmp/_src.csv
ncu -i $f --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin))
for h,u,v in zip(r[0],r[1],r[2]):
    if h in ('dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'): print('  ',h,u,v)
    if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
        try:
            if float(v)>0.2: print('   stall',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v)
        except: pass
"
