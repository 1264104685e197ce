#!/bin/bash
# Round 2: parity of the strip index + per-phase profile + bench on the three feature densities
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q -m gpu -x 2>&1 | tail -8
for w in kitti00 kitti00-ground kitti00-clustered; do timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --workload $w 2>gpurun_out/wl_$w.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['value']), d['phases_ms'], {k:v for k,v in d['config']['status_hist'].items() if v and k not in ('updated','second_dt','frames')})"; tail -2 gpurun_out/wl_$w.err; done
timeout 120 python scripts/phase_profile.py 592 2>&1 | python -c "
import json,sys
r=json.load(sys.stdin)
print('kernel_ms %.2f fps %.0f cycles/frame %.0f deferred %.1f fallback %.2f exact %.1f' % (r['kernel_ms'], r['fps'], r['cycles_per_frame_total'], r['n_deferred'], r['n_fallback'], r['n_exact']))
for k,v in r['phases'].items(): print('  %-18s %10d %.3f' % (k, int(v['cycles']), v['share']))"
