#!/bin/bash
# phase profile for prebuilt library variants gpurun_scratch/libmvosr_<tag>.so
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/libmvosr_default.so
for t in "$@"; do
echo "== variant $t"
cp gpurun_scratch/libmvosr_$t.so mvoscalerecovery_b200/csrc/libmvosr.so
timeout 120 python scripts/phase_profile.py 592 2>&1 | python -c "
import json,sys
r=json.load(sys.stdin)
print('kernel_ms %.2f fps %.0f deferred %.1f fallback %.2f' % (r['kernel_ms'], r['fps'], r['n_deferred'], r['n_fallback']))
print(' '.join('%s=%d' % (k, int(v['cycles'])) for k,v in r['phases'].items()))"
done
cp /tmp/libmvosr_default.so mvoscalerecovery_b200/csrc/libmvosr.so
