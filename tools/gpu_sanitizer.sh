#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (records, one-pass lists, leg masks, lean VJP, streaming); text to gpurun_out/.
mkdir -p gpurun_out
OUT=gpurun_out/compute_sanitizer.txt
echo "# compute-sanitizer on the round-2 kernels (B200, CUDA 12.9)" > $OUT
run() { echo >> $OUT; echo "\$ compute-sanitizer --tool $1 python -m pytest $2 -m gpu -q -x -k \"$3\"" >> $OUT
        timeout 1200 compute-sanitizer --tool $1 python -m pytest $2 -m gpu -q -x -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -8 >> $OUT; }
run memcheck  "tests/test_gpu_trace.py tests/test_gpu_stages.py tests/test_gpu_vjp.py tests/test_gpu_streaming.py tests/test_gpu_surface_grads.py" "ray_direction or culling_is_exact or binned or soft_sensors or cassegrain or leg_culling or two_stage or hard_hex or cylinder_caps or window or surface"
run racecheck "tests/test_gpu_trace.py tests/test_gpu_stages.py tests/test_gpu_vjp.py" "config1 or response_matrix_rows or ray_direction or binned or hard_hex or leg_culling"
run synccheck "tests/test_gpu_trace.py tests/test_gpu_stages.py" "config1 or response_matrix_rows or ray_direction or binned or leg_culling"
cat $OUT
