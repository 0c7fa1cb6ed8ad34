timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { python bench.py --workload $1 --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s %-14s %8.3f ms %7.2f Grays/s  cyl/ray %.3f' % ('$1', '$2', d['ms_per_step'], d['value']/1e9, d['roofline']['flops_per_ray']['mean_cylinders_tested']))"; }
run ct5_point_4096x115_hex unbinned
IACTRACE_B200_BIN_SAMPLES_MIN=64 run ct5_point_4096x115_hex binned64
run ct5_point_4096x4096_hex binned
run ct3_matrix_64x64_M1000 unbinned
IACTRACE_B200_BIN_OBSTRUCTIONS_MIN=10 run ct3_matrix_64x64_M1000 binned
IACTRACE_B200_BIN_OBSTRUCTIONS_MIN=10 IACTRACE_B200_BIN_SAMPLES_MIN=64 run ct3_matrix_64x64_M64 binned64
run cassegrain_1e9 default
