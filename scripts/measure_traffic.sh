#!/bin/bash
# The ncu evidence behind bench.py's roofline block (traffic, issue_frac, winstr_per_frame), taken from the bench command itself:
#   gpurun --timeout 900 -- 'bash scripts/measure_traffic.sh r02'
# then, in the build container:  python scripts/summarize_profiles.py gpurun_out/prof_r02_final.ncu-rep gpurun_out/launches_r02.csv r02
# (writes profiles/ncu_r02_*.txt, profiles/launches_r02*.{csv,txt} and profiles/traffic.json with the hash of the kernel sources).
RT=${1:-r02}
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$RT.csv \
    python bench.py --steps 2 --warmup 1 --cpu-sample 0 --dropin-frames 0 --no-densities > gpurun_out/ncu_launch_bench.log 2>&1
tail -2 gpurun_out/ncu_launch_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 3 -c 1 -o gpurun_out/prof_${RT}_final -f \
    python bench.py --steps 1 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --no-densities > gpurun_out/ncu_full_bench.log 2>&1
tail -2 gpurun_out/ncu_full_bench.log
