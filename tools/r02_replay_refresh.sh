#!/bin/bash
# Evidence refresh on ONE B200 after the replay-kernel change (warp-private tiles): GPU tests, the bench line, and the
# ncu capture of k_replay_paths.  Outputs under gpurun_out/.
set -x
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${R}_gpu_tests_1gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_gpu_tests_1gpu.log
tail -4 gpurun_out/${R}_gpu_tests_1gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err; echo "bench rc=$?"
MCTS_PROFILE_REPLAY=227328 timeout 300 python tools/profile_step.py --reps 2 > gpurun_out/${R}_replay_plain.log 2>&1; tail -2 gpurun_out/${R}_replay_plain.log
MCTS_PROFILE_REPLAY=227328 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_paths -s 1 -c 1 \
    -o gpurun_out/${R}_prof_replay_paths -f python tools/profile_step.py --reps 1 > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_c.log
