#!/bin/bash
# Round 2, call F: tests, config 5 A/B (soft-hex forward at 3 vs 4 resident blocks) + profiles as text.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -14
bash tools/gpu_ab.sh "ct5_cfg5_loss_grad_4096x115_softhex" 20
bash tools/gpu_profile_text.sh cfg5_vjp2 vjp_kernel 3 4.126e8 ct5_cfg5_loss_grad_4096x115_softhex
bash tools/gpu_profile_text.sh cfg5_fwd2 trace_kernel 4 4.126e8 ct5_cfg5_loss_grad_4096x115_softhex
ls -la gpurun_out | tail -6
