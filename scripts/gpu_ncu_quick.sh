#!/bin/bash
# one ncu --set full capture of the frame kernel on the bench workload (592 frames keep it short) + the parity tests
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 2 -c 1 -o gpurun_out/prof_${1:-quick} -f \
    python scripts/phase_profile.py 1184 > gpurun_out/ncu_quick.log 2>&1; tail -2 gpurun_out/ncu_quick.log
