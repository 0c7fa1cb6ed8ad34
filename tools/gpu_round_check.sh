#!/bin/bash
# Round-end style check: GPU tests, parity report, fmad invariance, bench (both arms), ncu launch list, text profiles.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/parity_report.py > gpurun_out/parity.json 2> gpurun_out/parity.err; echo "parity rc=$?"
timeout 900 python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 700 gpurun_out/bench_default.jsonl
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.jsonl 2>/dev/null; tail -c 900 gpurun_out/bench_reference.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b_ncu.log 2>&1
timeout 600 python tools/check_fmad_invariance.py > gpurun_out/fmad_invariance.log 2>&1; tail -2 gpurun_out/fmad_invariance.log
bash tools/gpu_profile_text.sh trace_v13 trace_kernel 3 4.126e8 ct5_point_4096x115_hex


ls -la gpurun_out | tail -8
