#!/bin/bash
mkdir -p gpurun_out


run() { python bench.py --workload $1 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$2', d['ms_per_step'], 'ms', d['value']/1e9, 'Grays/s')"; }
for w in ct5_point_4096x115_hex ct5_point_4096x115_square ct3_matrix_64x64_M64 ct3_matrix_64x64_M1000 ct5_point_4096x4096_hex cassegrain_1e9; do run $w; done
python tools/gpu_cfg5.py
