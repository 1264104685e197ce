#!/bin/bash
# star counters for prebuilt counter variants gpurun_scratch/libmvosr_<tag>.so
cp mvoscalerecovery_b200/csrc/libmvosr.so /tmp/libmvosr_default.so
for t in "$@"; do
echo "== variant $t"
cp gpurun_scratch/libmvosr_$t.so mvoscalerecovery_b200/csrc/libmvosr.so
timeout 120 python scripts/star_counters.py 592 2>&1 | tail -3
done
cp /tmp/libmvosr_default.so mvoscalerecovery_b200/csrc/libmvosr.so
