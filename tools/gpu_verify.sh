#!/bin/bash
# Round-end style check: GPU tests, default bench (both arms), ncu launch list, one full capture of trace_kernel.
TAG=${1:-v6}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.jsonl
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.jsonl 2>/dev/null; tail -1 gpurun_out/bench_reference.jsonl
bash tools/gpu_run1.sh 2>&1 | grep -v pytest > gpurun_out/workloads.log; cat gpurun_out/workloads.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
