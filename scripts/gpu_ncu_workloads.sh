#!/bin/bash
# one ncu --set full capture of the frame kernel on each of the other workloads (perspective, clustered, dense), summarised on the box
# (the reports are too large to bring back together): gpurun_out/ncu_r02_frame_kernel_<workload>.txt
WANT='gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__inst_executed.avg.per_cycle_elapsed|smsp__inst_executed.sum|smsp__issue_active.avg.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed.ratio|sm__warps_active.avg.pct_of_peak_sustained_active|launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active|pcsamp_warps_issue_stalled_(wait|barrier|long_scoreboard|no_instructions|not_selected|math_pipe_throttle|short_scoreboard)$'
for wl in kitti00-ground kitti00-clustered dense; do
timeout 500 ncu --set full --clock-control none -k regex:frame_kernel -s 3 -c 1 -o /tmp/prof_$wl -f \
    python bench.py --workload $wl --steps 1 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --no-densities > /tmp/ncu_$wl.log 2>&1
ncu -i /tmp/prof_$wl.ncu-rep --page raw --csv 2>/dev/null | WANT="$WANT" WL=$wl python -c "
import csv, os, re, sys
rows = list(csv.reader(sys.stdin)); h, u, v = rows[0], rows[1], rows[2]
print('ncu --set full --clock-control none -k regex:frame_kernel -s 3 -c 1 python bench.py --workload %s --steps 1 --warmup 3 ...' % os.environ['WL'])
for i, n in enumerate(h):
    if re.search(os.environ['WANT'], n): print('%-72s %-16s %s' % (n, u[i], v[i]))
" > gpurun_out/ncu_r02_frame_kernel_$wl.txt
tail -3 gpurun_out/ncu_r02_frame_kernel_$wl.txt | cut -c1-120
done
