#!/bin/bash
# usage: scripts/ncu_summary.sh gpurun_out/prof_X.ncu-rep > profiles/X.txt   (run in the build container, no GPU needed)
rep=$1
echo "# ncu summary of $(basename $rep) (ncu --set full --clock-control none --import-source on)"
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
r=list(csv.reader(sys.stdin)); hdr=r[0]; unit=r[1]; vals=r[2]
want=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed',
'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg.per_second','sm__cycles_elapsed.avg','smsp__inst_executed.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__inst_executed_pipe_uniform','smsp__warp_issue_stalled']
for h,u,v in zip(hdr,unit,vals):
    if h in want: print(f'{h:75s} {u:18s} {v}')
"
ncu -i $rep --page source --csv 2>/dev/null > /tmp/_src.csv
python3 $(dirname $0)/ncu_top.py /tmp/_src.csv 12
