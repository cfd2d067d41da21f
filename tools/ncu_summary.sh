#!/bin/bash
# usage: tools/ncu_summary.sh report.ncu-rep > profiles/<name>.md   -- compact, committed summary of an ncu --set full capture
REP=$1
echo "# ncu summary of $(basename $REP)"
echo
echo '```'
ncu -i $REP --page details 2>/dev/null | grep -E "popoa_kernel|pwfa_kernel|chain_kernel|Duration|SM Frequency|Elapsed Cycles|Executed Ipc|Issue Slots Busy|Issued Warp|No Eligible|Eligible Warps|Warp Cycles Per Issued|Avg. Active Threads|Avg. Not Predicated|Registers Per Thread|Dynamic Shared|Theoretical Occupancy|Achieved Occupancy|DRAM Throughput|Mem Busy|L1/TEX Hit|L2 Hit|Grid Size|Block Size|highest-utilized"
echo '```'
echo
echo "## raw metrics"
echo '```'
ncu -i $REP --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
want=('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__inst_executed.sum','smsp__inst_executed.sum','sm__pipe_alu_cycles_active','sm__inst_executed_pipe_alu','sm__inst_executed_pipe_fma','smsp__thread_inst_executed.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct')
for k,x in zip(h,v):
    if any(k.startswith(w) for w in want): print(k,x)
print('--- warp stall reasons (warp-cycles per issued instruction) ---')
st=[(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),x) for k,x in zip(h,v) if 'smsp__average_warp' in k and 'issue_stalled' in k and 'ratio' in k and 'not_issued' not in k]
for k,x in sorted(st,key=lambda t:-float(t[1]))[:10]: print(k,x)
"
echo '```'
echo
echo "## hottest source lines (share of warp instructions, avg active threads, share of stall samples)"
echo '```'
python3 $(dirname $0)/ncu_lines.py $REP ${2:-popoa_kernelILi3} 25
echo '```'
