#!/bin/bash
# Round 2, call E: tests (surface gradients new), fmad invariance, config 5 / Cassegrain profiles as text, bench line.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -14
timeout 600 python tools/check_fmad_invariance.py > gpurun_out/fmad_invariance.log 2>&1; echo "fmad rc=$?"; tail -6 gpurun_out/fmad_invariance.log
bash tools/gpu_profile_text.sh cfg5_vjp vjp_kernel 3 4.126e8 ct5_cfg5_loss_grad_4096x115_softhex
bash tools/gpu_profile_text.sh cfg5_fwd trace_kernel 4 4.126e8 ct5_cfg5_loss_grad_4096x115_softhex
bash tools/gpu_profile_text.sh cass trace_kernel 3 1.00002e9 cassegrain_1e9
timeout 900 python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.jsonl
ls -la gpurun_out | tail -12
