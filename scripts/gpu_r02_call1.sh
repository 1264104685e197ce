#!/bin/bash
# Round 2, first GPU call: the whole GPU suite, the N1 measurements (primitives incl. find_essential rows with OpenCV beside them,
# one ncu capture of find_essential_kernel, compute-sanitizer over the small workload) and a baseline bench line of the round-1 build.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_r02_call1.sh'
mkdir -p gpurun_out
echo "== whole GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "== bench (round-1 build, baseline of this pool)"; timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 50 > gpurun_out/bench_r02_base_n1.json 2> gpurun_out/bench_r02_base.err; tail -c 600 gpurun_out/bench_r02_base_n1.json
echo "== primitives"; timeout 600 python scripts/bench_primitives.py > gpurun_out/primitives_r02.json 2> gpurun_out/primitives_r02.err; tail -45 gpurun_out/primitives_r02.json
echo "== ncu find_essential"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:find_essential -c 1 -o gpurun_out/prof_r02_find_essential -f python scripts/sanitize_small.py > gpurun_out/ncu_r02_find_essential.log 2>&1; tail -3 gpurun_out/ncu_r02_find_essential.log
echo "== compute-sanitizer"
for tool in memcheck racecheck; do
    timeout 420 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_r02_$tool.txt 2>&1; tail -4 gpurun_out/sanitizer_r02_$tool.txt
done
