#!/bin/bash
# Round-2 multi-GPU evidence (N B200s, N = number visible; gpurun --gpus N -- "bash tools/r02_multi_gpu.sh [grouptests|notests]"): GPU tests incl. the multi-device group on every GPU, the N-GPU bench
# line (torchrun), BASELINE's five configs at full size over N ranks, and the same through ONE process (MCTS_DEVICES=all).
set -x
R=r02
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
if [ "$1" == "grouptests" ]; then   # only the multi-device tests (an 8-GPU box is charged eight-fold)
  timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s -k "group or device_list or nccl" > gpurun_out/${R}_tests_${N}gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_tests_${N}gpu.log
  tail -6 gpurun_out/${R}_tests_${N}gpu.log
elif [ "$1" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${R}_tests_${N}gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_tests_${N}gpu.log
  tail -6 gpurun_out/${R}_tests_${N}gpu.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 \
    > gpurun_out/${R}_bench_${N}gpu.json 2> gpurun_out/${R}_bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/${R}_bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/full_configs.py \
    > gpurun_out/${R}_full_configs_${N}gpu.txt 2> gpurun_out/${R}_full_configs_${N}gpu.err; echo "full_configs rc=$?"
cat gpurun_out/${R}_full_configs_${N}gpu.txt
MCTS_DEVICES=all timeout 600 python tools/group_configs.py > gpurun_out/${R}_group_configs_${N}gpu.txt 2>&1; echo "group rc=$?"
cat gpurun_out/${R}_group_configs_${N}gpu.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${R}_bench_${N}gpu.json"))
for k in ("value","ms_per_step","e2e","partition_sha","inlib_multi_gpu","strong_scaling","c5_full"):
    print(k, json.dumps(d.get(k))[:400])
PY
