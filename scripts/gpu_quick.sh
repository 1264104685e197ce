#!/bin/bash
# parity tests + phase profile in one gpurun call
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -${2:-15}
timeout 120 python scripts/phase_profile.py ${1:-592} 2>&1 | python -c "
import json,sys
r=json.load(sys.stdin)
print('kernel_ms %.2f fps %.0f cycles/frame %.0f deferred %.1f fallback %.2f exact %.1f' % (r['kernel_ms'], r['fps'], r['cycles_per_frame_total'], r['n_deferred'], r['n_fallback'], r['n_exact']))
for k,v in r['phases'].items(): print('  %-18s %10d %.3f' % (k, int(v['cycles']), v['share']))"
