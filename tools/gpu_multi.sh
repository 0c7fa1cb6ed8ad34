#!/bin/bash
# Multi-GPU call (gpurun --gpus N): NCCL correctness tests on 2 GPUs, then the bench under torchrun at N ranks
# (weak headline + the `strong` block: fixed-size configs 2, 3, 4 split N ways).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/pytest_gpu_multi_${N}gpu.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_gpu_multi_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n${N}.jsonl 2> gpurun_out/bench_n${N}.err; echo "bench N=$N rc=$?"
tail -c 3000 gpurun_out/bench_n${N}.jsonl; tail -5 gpurun_out/bench_n${N}.err
