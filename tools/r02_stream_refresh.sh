#!/bin/bash
# Evidence refresh after the base-stream change (eight draws per Philox block): everything whose summary hashes depend on the
# generator stream.  ONE B200: GPU tests, bench line, full-size configs (ranks and in-library group).
# N > 1 (gpurun --gpus N): multi-device tests, N-GPU bench line, full-size configs.
set -x
R=r02
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
if [ "$N" == "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${R}_gpu_tests_1gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_gpu_tests_1gpu.log
  tail -4 gpurun_out/${R}_gpu_tests_1gpu.log
  timeout 900 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err; echo "bench rc=$?"
  timeout 600 python tools/full_configs.py > gpurun_out/${R}_full_configs_1gpu.txt 2>&1; cat gpurun_out/${R}_full_configs_1gpu.txt
else
  timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "group or device_list or nccl" > gpurun_out/${R}_gpu_tests_${N}gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_gpu_tests_${N}gpu.log
  tail -4 gpurun_out/${R}_gpu_tests_${N}gpu.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
      > gpurun_out/${R}_bench_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; echo "bench rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/full_configs.py \
      > gpurun_out/${R}_full_configs_${N}gpu.txt 2> gpurun_out/${R}_full_configs_${N}gpu.err; echo "full_configs rc=$?"
  cat gpurun_out/${R}_full_configs_${N}gpu.txt
fi
MCTS_DEVICES=all timeout 600 python tools/group_configs.py > gpurun_out/${R}_group_configs_${N}gpu.txt 2>&1; echo "group rc=$?"
cat gpurun_out/${R}_group_configs_${N}gpu.txt
