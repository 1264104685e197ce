#!/bin/bash
# Data-race check of find_essential_kernel without a GPU: the kernel SOURCE runs on the host with one OS thread per CUDA thread
# (tests/host_sim/fp5_kernel_emu.cpp: __syncthreads = pthread barrier, __shared__ = static storage) under ThreadSanitizer.
# A missing or misplaced barrier shows up as a data race on the shared arrays (checked: with the barriers disabled TSan
# reports the races).  Usage: scripts/tsan_five_point_kernel.sh  -> prints the per-frame results and the number of TSan warnings.
set -e
cd "$(dirname "$0")/../tests/host_sim"
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -Wno-unknown-pragmas -o /tmp/fp5_tsan fp5_kernel_emu.cpp fp5_tsan_main.cpp
/tmp/fp5_tsan 2>/tmp/fp5_tsan.err
echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' /tmp/fp5_tsan.err || true)"
# the same for bucket_kernel (bitonic sort in shared memory, survivors compacted in order)
g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -Wno-unknown-pragmas -o /tmp/bucket_tsan bucket_kernel_emu.cpp bucket_tsan_main.cpp
/tmp/bucket_tsan 2>/tmp/bucket_tsan.err
echo "ThreadSanitizer warnings (bucket_kernel): $(grep -c 'WARNING: ThreadSanitizer' /tmp/bucket_tsan.err || true)"
