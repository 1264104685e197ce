#!/bin/bash
# Compiles the product's thread-path star code (csrc/star.cuh, csrc/predicates.cuh) for the HOST so
# the CPU test-suite can execute the very same source.  Test infrastructure.
set -e
cd "$(dirname "$0")"
g++ -O2 -std=c++17 -shared -fPIC -x c++ -o libhostsim.so star_host.cu
