#!/bin/bash
# round-1 session-2 GPU check: stage-path parity, Cassegrain timings (2 vs 3 resident blocks), ncu capture
mkdir -p gpurun_out
python -m pytest tests/test_gpu_stages.py tests/test_gpu_trace.py tests/test_gpu_vjp.py -x -q -m gpu > gpurun_out/pytest_stages.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_stages.log
python bench.py --workload cassegrain_1e9 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cass_a.log 2>gpurun_out/cass_a.err; tail -1 gpurun_out/cass_a.log | cut -c1-400
IACTRACE_B200_LIB=$PWD/variants/libiactrace_b200_st3.so python bench.py --workload cassegrain_1e9 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/cass_b.log 2>gpurun_out/cass_b.err; tail -1 gpurun_out/cass_b.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -c 1 -f -o gpurun_out/prof_cass python bench.py --workload cassegrain_1e9 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cass.log 2>&1; echo "ncu rc=$?"
