#!/bin/bash
# ncu captures of the other workloads' kernels, summarised on the box (the reports are too large to bring back),
# + one bench line per workload (profiles/ evidence).
mkdir -p gpurun_out
for w in cassegrain_1e9 ct3_matrix_64x64_M1000; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o /tmp/prof_$w python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/ncu_$w.log 2>&1
  python tools/summarize_ncu.py /tmp/prof_$w.ncu-rep gpurun_out/prof_$w.txt > /dev/null 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"vjp_kernel|trace_kernel" -s 4 -c 2 -f -o /tmp/prof_cfg5 python tools/gpu_vjp_profile.py > gpurun_out/cfg5_ncu.log 2>&1
python tools/summarize_ncu.py /tmp/prof_cfg5.ncu-rep gpurun_out/prof_cfg5.txt > /dev/null 2>&1
rm -f gpurun_out/bench_workloads.jsonl
for w in ct5_point_4096x115_square ct3_matrix_64x64_M64 ct3_matrix_64x64_M1000 ct3_matrix_512x512_M64 ct5_point_4096x4096_hex cassegrain_1e9; do
  timeout 300 python bench.py --workload $w --steps 30 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/bench_workloads.jsonl
done
wc -l gpurun_out/bench_workloads.jsonl; ls -la gpurun_out
