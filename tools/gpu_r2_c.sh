#!/bin/bash
# Round 2, call C: fmad-invariance (render.cu only), GPU tests, A/B of records x occupancy x second fast path,
# ncu of the product build, streaming workload, bench line.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python tools/check_fmad_invariance.py > gpurun_out/fmad_invariance.log 2>&1; echo "fmad rc=$?"; tail -4 gpurun_out/fmad_invariance.log
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -12
bash tools/gpu_ab.sh "ct5_point_4096x115_hex ct5_point_4096x115_square ct3_matrix_64x64_M64 ct3_matrix_64x64_M1000 cassegrain_1e9" 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu2.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_default.jsonl
for wb in 50331648 1099511627776; do IACTRACE_B200_WINDOW_TABLE_BYTES=$wb timeout 300 python bench.py --workload ct5_point_4096x4096_hex --steps 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('M4096 window_bytes=$wb', d['ms_per_step'], d['value'])"; done
ls -la gpurun_out | tail -5
