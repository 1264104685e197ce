#!/bin/bash
# A/B of prebuilt library variants gpurun_scratch/libmvosr_<tag>.so against the default build: kernel-only frames/s on the three feature densities
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/libmvosr_default.so
for t in default "$@"; do
[ "$t" != default ] && cp gpurun_scratch/libmvosr_$t.so mvoscalerecovery_b200/csrc/libmvosr.so
for wl in kitti00 kitti00-ground kitti00-clustered; do
timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --no-densities 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$t $wl', round(j['value']), round(j['e2e']['value']), j['roofline']['kernel_ms'], j['config'].get('status_hist'))"
done; done
cp /tmp/libmvosr_default.so mvoscalerecovery_b200/csrc/libmvosr.so
