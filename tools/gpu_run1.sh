#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
for v in "" st2 st4; do
  if [ -n "$v" ]; then export IACTRACE_B200_LIB=$PWD/variants/libiactrace_b200_$v.so; fi
  python bench.py --workload cassegrain_1e9 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cass_$v.log 2>gpurun_out/cass_$v.err; echo "variant=$v $(tail -1 gpurun_out/cass_$v.log | cut -c1-200)"
done
