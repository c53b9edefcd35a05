#!/bin/bash
# Round-2 final evidence on ONE B200 (run under gpurun): GPU tests, bench line, reference arm, ncu launch list and full captures
# of the hot kernel, the replay kernel and the curve-export kernel.  Outputs under gpurun_out/.
set -x
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${R}_gpu_tests_1gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_gpu_tests_1gpu.log
tail -6 gpurun_out/${R}_gpu_tests_1gpu.log
timeout 900 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
timeout 600 python tools/full_configs.py > gpurun_out/${R}_full_configs_1gpu.txt 2>&1; cat gpurun_out/${R}_full_configs_1gpu.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $CMD > gpurun_out/plain_a.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sim_paths -s 12 -c 1 \
    -o gpurun_out/${R}_prof_bench_sim_paths -f $CMD > gpurun_out/ncu_b.log 2>&1
MCTS_PROFILE_REPLAY=227328 timeout 300 python tools/profile_step.py --reps 2 > gpurun_out/${R}_replay_plain.log 2>&1; tail -2 gpurun_out/${R}_replay_plain.log
MCTS_PROFILE_REPLAY=227328 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_replay_paths -s 1 -c 1 \
    -o gpurun_out/${R}_prof_replay_paths -f python tools/profile_step.py --reps 1 > gpurun_out/ncu_c.log 2>&1
timeout 300 python tools/curve_step.py > gpurun_out/${R}_curve_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:k_sim_paths -s 1 -c 1 -o gpurun_out/${R}_prof_curve_export -f python tools/curve_step.py > gpurun_out/ncu_d.log 2>&1
tail -2 gpurun_out/${R}_curve_plain.log
timeout 300 python tools/profile_step.py --reps 3 --dedup > gpurun_out/${R}_dedup_plain.log 2>&1; tail -1 gpurun_out/${R}_dedup_plain.log
ls -la gpurun_out | tail -25
