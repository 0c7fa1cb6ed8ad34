#!/bin/bash
# Round 2, call D: tests, fmad invariance, config 5 (lean VJP, cotangent cache, two-slot soft-hex forward) A/B + ncu,
# Cassegrain / matrix re-check and ncu of the stage kernel.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -12
timeout 600 python tools/check_fmad_invariance.py > gpurun_out/fmad_invariance.log 2>&1; echo "fmad rc=$?"; tail -12 gpurun_out/fmad_invariance.log
bash tools/gpu_ab.sh "ct5_cfg5_loss_grad_4096x115_softhex cassegrain_1e9 ct3_matrix_64x64_M64 ct5_point_4096x115_hex" 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vjp_kernel|trace_kernel" -s 6 -c 2 -f -o gpurun_out/prof_cfg5_${TAG} python bench.py --workload ct5_cfg5_loss_grad_4096x115_softhex --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 3 -c 1 -f -o gpurun_out/prof_cass_${TAG} python bench.py --workload cassegrain_1e9 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu4.log 2>&1
timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; echo "parity rc=$?"; tail -2 gpurun_out/parity.err | cut -c1-300
ls -la gpurun_out | tail -6
