#!/bin/bash
# frames/s and features/s of the fused kernel against the number of correspondences per frame (shared-memory staging up to 3 264 ROI
# features, per-CTA global-memory slabs beyond): the per-feature cost across the two staging modes
for n in 600 1250 2500 4000 5000 8000 12500 25000; do
fr=$(( 2500 * 4541 / n )); [ $fr -gt 4541 ] && fr=4541; [ $fr -lt 296 ] && fr=296
timeout 300 python bench.py --features $n --frames $fr --steps 5 --warmup 3 --cpu-sample 0 --dropin-frames 0 --no-fleet --no-densities 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); c=j['config']; n=$n; fps=j['value']
print(json.dumps({'correspondences_per_frame': n, 'frames': c['frames_total'], 'frames_per_s': fps, 'correspondences_per_s': fps*n, 'e2e_frames_per_s': j['e2e']['value'], 'kernel_ms': j['roofline']['kernel_ms'], 'status_hist': c.get('status_hist')}))"
done
