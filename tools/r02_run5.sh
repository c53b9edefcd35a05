#!/bin/bash
# Round-2 GPU session 5 (ONE B200): first-wave staggering experiment.
set -x
mkdir -p gpurun_out
for st in 0 1; do
  MCTS_STAGGER=$st timeout 300 python tools/profile_step.py --reps 5 > gpurun_out/r02_stagger_${st}.log 2>&1; echo "== stagger $st"; tail -2 gpurun_out/r02_stagger_${st}.log
done
for pct in 50 150 200 300; do
  MCTS_STAGGER_PCT=$pct timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/r02_stagger_pct${pct}.log 2>&1; echo "== pct $pct"; tail -1 gpurun_out/r02_stagger_pct${pct}.log
done
MCTS_STAGGER=1 timeout 300 python tools/profile_step.py --reps 4 --dedup > gpurun_out/r02_stagger_dedup1.log 2>&1; tail -1 gpurun_out/r02_stagger_dedup1.log
MCTS_STAGGER=0 timeout 300 python tools/profile_step.py --reps 4 --dedup > gpurun_out/r02_stagger_dedup0.log 2>&1; tail -1 gpurun_out/r02_stagger_dedup0.log
MCTS_STAGGER=1 timeout 300 python tools/profile_step.py --reps 4 --mode markov2 --no-int > gpurun_out/r02_stagger_m2_1.log 2>&1; tail -1 gpurun_out/r02_stagger_m2_1.log
MCTS_STAGGER=0 timeout 300 python tools/profile_step.py --reps 4 --mode markov2 --no-int > gpurun_out/r02_stagger_m2_0.log 2>&1; tail -1 gpurun_out/r02_stagger_m2_0.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_stagger_tests.log 2>&1; tail -3 gpurun_out/r02_stagger_tests.log
