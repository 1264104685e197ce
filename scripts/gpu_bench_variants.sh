#!/bin/bash
# bench.py for prebuilt library variants gpurun_scratch/libmvosr_<tag>.so (and the default build)
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/libmvosr_default.so
for t in default "$@"; do
[ "$t" != default ] && cp gpurun_scratch/libmvosr_$t.so mvoscalerecovery_b200/csrc/libmvosr.so
python bench.py --steps 10 --warmup 3 --cpu-sample 0 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$t', round(j['value']), round(j['e2e']['value']), j['roofline']['kernel_ms'])"
done
cp /tmp/libmvosr_default.so mvoscalerecovery_b200/csrc/libmvosr.so
