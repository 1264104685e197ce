#!/bin/bash
# First GPU call of the next round: everything the five-point RANSAC kernel (mvosr_find_essential_frames) still owes, in one gpurun.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
# It was written and CPU-verified (host build + pthread emulation + ThreadSanitizer) after the round's GPU minutes had run out.
mkdir -p gpurun_out
echo "== parity (new kernel)"; timeout 600 python -m pytest tests/test_gpu_zz_essential.py tests/test_gpu_zzz_bucket.py -q -m gpu 2>&1 | tail -15
echo "== whole GPU suite"; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
echo "== compute-sanitizer memcheck / racecheck (small)"
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_r02_$tool.txt 2>&1; tail -3 gpurun_out/sanitizer_r02_$tool.txt
done
echo "== primitives (find_essential rows + OpenCV on the host)"; timeout 900 python scripts/bench_primitives.py > gpurun_out/primitives_r02.json 2> gpurun_out/primitives_r02.err; tail -30 gpurun_out/primitives_r02.json
echo "== ncu: the find_essential launches of the sanitizer workload"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:find_essential -c 1 -o gpurun_out/prof_r02_find_essential -f python scripts/sanitize_small.py > gpurun_out/ncu_r02_find_essential.log 2>&1; tail -3 gpurun_out/ncu_r02_find_essential.log
