#!/bin/bash
# Build variants of the trace kernel with different occupancy targets and time the main workloads.
for mb in 2 3 4 5; do
  IACT_NVCC_EXTRA="-DIACT_MIN_BLOCKS=$mb" python iactrace_b200/csrc/build.py --force > /dev/null 2>&1
  echo "=== IACT_MIN_BLOCKS=$mb"
  python tools/gpu_explore.py 2>&1 | grep -E "CT5 render|CT3 response|CT3 render" | grep -v "cull=False"
done
python iactrace_b200/csrc/build.py --force > /dev/null 2>&1
