#!/bin/bash
# Round 2, call A: GPU test-suite, parity report, default bench (both arms, with the workloads block), A/B of the
# float64 square-camera accumulation against the float32 per-ray red.global it replaces.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; echo "parity rc=$?"; tail -3 gpurun_out/parity.err
timeout 600 python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_default.jsonl
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.jsonl 2>/dev/null; tail -c 600 gpurun_out/bench_reference.jsonl
bash tools/gpu_ab.sh "ct5_point_4096x115_square cassegrain_1e9 ct5_point_4096x115_hex" 20
