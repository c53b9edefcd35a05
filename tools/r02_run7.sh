#!/bin/bash
set -x
mkdir -p gpurun_out
for v in perm1203 perm2103; do MCTS_LIB=tools/variants/libmcts_$v.so timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/r02_x_$v.log 2>&1; echo "== $v"; tail -1 gpurun_out/r02_x_$v.log; done
MCTS_LIB=tools/variants/libmcts_b192.so timeout 300 python tools/profile_step.py --reps 4 --block 192 > gpurun_out/r02_x_b192.log 2>&1; echo "== b192"; tail -1 gpurun_out/r02_x_b192.log
MCTS_LIB=tools/variants/libmcts_b192.so timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/r02_x_b192_128.log 2>&1; echo "== b192 lib at 128"; tail -1 gpurun_out/r02_x_b192_128.log
timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/r02_x_shipped.log 2>&1; echo "== shipped"; tail -1 gpurun_out/r02_x_shipped.log
