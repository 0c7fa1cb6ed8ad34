#!/bin/bash
# compute-sanitizer passes over the kernels (memcheck on a broad subset, racecheck + synccheck on the shared-memory paths).
mkdir -p gpurun_out
{
echo "# compute-sanitizer on the round-1 kernels with the work queues (B200, CUDA 12.9)"
echo
echo '$ compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_trace.py tests/test_gpu_stages.py tests/test_gpu_vjp.py -m gpu -q -x -k "edge or culling_is_exact or binned or soft_sensors or cassegrain or two_stage or hard_hex or cylinder_caps"'
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_trace.py tests/test_gpu_stages.py tests/test_gpu_vjp.py -m gpu -q -x -k "edge or culling_is_exact or binned or soft_sensors or cassegrain or two_stage or hard_hex or cylinder_caps" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -20
echo
echo '$ compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_trace.py tests/test_gpu_vjp.py -m gpu -q -x -k "config1 or response_matrix_rows or many_sources or binned or hard_hex"'
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_trace.py tests/test_gpu_vjp.py -m gpu -q -x -k "config1 or response_matrix_rows or many_sources or binned or hard_hex" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|Race" | head -20
echo
echo '$ compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_trace.py -m gpu -q -x -k "config1 or response_matrix_rows or many_sources or binned"'
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_trace.py -m gpu -q -x -k "config1 or response_matrix_rows or many_sources or binned" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Barrier|error" | head -20
} > gpurun_out/sanitizer.txt 2>&1
cat gpurun_out/sanitizer.txt
