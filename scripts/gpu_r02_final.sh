#!/bin/bash
# Round 2, final single-GPU evidence run on the committed build:
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r02_final.sh'
# whole GPU suite, the bench line (with cpu_baseline = the unmodified reference, drop-in leg, fleet block, other densities), the reference
# arm, the ncu launch list + full capture of the bench command, the dense and fleet workloads, the RANSAC sweep, compute-sanitizer.
mkdir -p gpurun_out
echo "== whole GPU suite"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
echo "== bench N=1"; timeout 600 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; tail -c 400 gpurun_out/bench_r02_final_n1.json
echo "== reference arm"; timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; tail -c 600 gpurun_out/bench_r02_reference.json
echo "== ncu"; bash scripts/measure_traffic.sh r02 2>&1 | tail -c 300
echo "== dense"; timeout 600 python bench.py --workload dense --cpu-sample 4 > gpurun_out/bench_r02_dense_n1.json 2> gpurun_out/bench_r02_dense.err; tail -c 300 gpurun_out/bench_r02_dense_n1.json
echo "== fleet"; timeout 600 python bench.py --workload fleet --cpu-sample 0 --dropin-frames 0 > gpurun_out/bench_r02_fleet_n1.json 2> gpurun_out/bench_r02_fleet.err; tail -c 300 gpurun_out/bench_r02_fleet_n1.json
echo "== ransac sweep"; timeout 600 python scripts/ransac_sweep.py > gpurun_out/ransac_sweep_r02.json 2> gpurun_out/ransac_sweep_r02.err; tail -c 200 gpurun_out/ransac_sweep_r02.json
echo "== compute-sanitizer"
for tool in memcheck racecheck; do
    timeout 420 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_r02_$tool.txt 2>&1; tail -3 gpurun_out/sanitizer_r02_$tool.txt
done
