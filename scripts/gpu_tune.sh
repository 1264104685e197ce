#!/bin/bash
# sweep of the strip index knobs (k = points per square cell, window factor) on the three feature densities: kernel ms of 592 frames
for data in uniform ground clustered; do
for k in ${KS:-1.3 1.5 1.8}; do for w in ${WS:-2.2 2.5 2.8}; do
MVOSR_DATA=$data MVOSR_DENSITY=$k MVOSR_WFAC=$w timeout 120 python scripts/phase_profile.py 592 2>&1 | python -c "
import json,sys
r=json.load(sys.stdin); p=r['phases']
print('$data k=$k w=$w ms %.3f grid %d pair1 %d wrap1 %d pair2 %d wrap2 %d nwrap %d l3 %.1f' % (r['kernel_ms'], p['grid1']['cycles'], p['stars1_pair']['cycles'], p['stars1']['cycles']-p['stars1_pair']['cycles'], p['stars2_pair']['cycles'], p['stars2']['cycles']-p['stars2_pair']['cycles'], p['n_to_wrap_path']['cycles'], r['n_deferred']))"
done; done; done
