#!/bin/bash
# Round-2 GPU session 2 (TWO B200s): GPU tests incl. the multi-device group, replay kernel rate, 2-GPU bench line.
set -x
R=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/${R}_smi2.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${R}_tests_2gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_tests_2gpu.log
tail -8 gpurun_out/${R}_tests_2gpu.log
MCTS_PROFILE_REPLAY=200000 timeout 600 python tools/profile_step.py --reps 2 > gpurun_out/${R}_replay_rate.log 2>&1; tail -3 gpurun_out/${R}_replay_rate.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/${R}_bench_2gpu.json 2> gpurun_out/${R}_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/${R}_bench_2gpu.err
head -c 1500 gpurun_out/${R}_bench_2gpu.json
