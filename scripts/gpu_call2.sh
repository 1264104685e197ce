#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -3
bash scripts/gpu_ab.sh head
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/keep.so
cp gpurun_scratch/libmvosr_wc.so mvoscalerecovery_b200/csrc/libmvosr.so
for d in uniform ground; do echo "== $d"; MVOSR_WRAP=1 MVOSR_DATA=$d timeout 120 python scripts/star_counters.py 592; done
cp /tmp/keep.so mvoscalerecovery_b200/csrc/libmvosr.so
for dens in 1.2 1.5 1.8; do for wf in 2.2 2.5 2.8; do
MVOSR_DENSITY=$dens MVOSR_WFAC=$wf timeout 120 python scripts/phase_profile.py 592 2>&1 | python -c "
import json,sys
r=json.load(sys.stdin); p=r['phases']
print('uniform dens=$dens wfac=$wf ms %.3f pair1 %d wrap1 %d pair2 %d wrap2 %d nwrap %d l3 %.1f' % (r['kernel_ms'], p['stars1_pair']['cycles'], p['stars1']['cycles']-p['stars1_pair']['cycles'], p['stars2_pair']['cycles'], p['stars2']['cycles']-p['stars2_pair']['cycles'], p['n_to_wrap_path']['cycles'], r['n_deferred']))"
done; done
