#!/bin/bash
# profiles/<tag>_ncu_<cfg>_summary.txt from gpurun_out/<tag>_<cfg>.ncu-rep (run here, no GPU needed)
tag=$1; cfg=$2; note=$3
rep=gpurun_out/${tag}_${cfg}.ncu-rep
out=profiles/${tag}_ncu_${cfg}_summary.txt
M='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__inst_executed.sum$|sm__inst_executed_pipe_(fma|fp64|alu|lsu|xu).avg.pct|sm__pipe_tensor_cycles_active.avg.pct|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|waves_per_multiprocessor|occupancy_limit_registers|occupancy_limit_shared_mem)|l1tex__throughput.avg.pct|lts__throughput.avg.pct|smsp__warps_eligible.avg.per_cycle_active'
echo "# ncu --set full --clock-control none --import-source on (${note}); a number under the profiler is not a bench value" > $out
ncu -i $rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys,re
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
print('# kernel:', vals[hdr.index('Kernel Name')])
pat=re.compile(r'$M')
for h,u,v in zip(hdr,units,vals):
    if pat.search(h): print(f'{h:90s} {v:>18s} {u}')
" >> $out
cat $out
