#!/bin/bash
# like gpu_ab.sh, and the parity + property tests on every variant first (a hang is cut by timeout)
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/libmvosr_default.so
for t in default "$@"; do
[ "$t" != default ] && cp gpurun_scratch/libmvosr_$t.so mvoscalerecovery_b200/csrc/libmvosr.so
echo "== $t: $(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -1)"
for wl in kitti00 kitti00-ground kitti00-clustered; do
timeout 120 python bench.py --workload $wl --steps 5 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --no-densities 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$t $wl', round(j['value']), round(j['e2e']['value']), j['roofline']['kernel_ms'], j['config'].get('status_hist',{}).get('updated'))"
done; done
cp /tmp/libmvosr_default.so mvoscalerecovery_b200/csrc/libmvosr.so
