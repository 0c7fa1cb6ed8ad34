#!/bin/bash
# Round 2, call B: fmad-invariance of the explicit per-ray chain, GPU tests, parity report, A/B of the kernel-diet
# switches, ncu launch list + full capture of trace_kernel on the default workload.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 600 python tools/check_fmad_invariance.py > gpurun_out/fmad_invariance.log 2>&1; echo "fmad rc=$?"; tail -4 gpurun_out/fmad_invariance.log
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -12
timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; echo "parity rc=$?"; tail -2 gpurun_out/parity.err | cut -c1-400
bash tools/gpu_ab.sh "ct5_point_4096x115_hex ct5_point_4096x115_square ct3_matrix_64x64_M64 ct3_matrix_64x64_M1000 cassegrain_1e9" 20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_${TAG} python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -8
