#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): bench line, reference arm, ncu launch list + full
# capture of the bench command, and the HBM-bound curve-export capture.  Outputs under gpurun_out/.
set -x
R=${1:-r01}
python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_a.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
$CMD > gpurun_out/plain_b.json 2>/dev/null && ncu --set full --clock-control none --import-source on -k regex:k_sim_paths -s 12 -c 1 \
    -o gpurun_out/${R}_prof_bench_sim_paths -f $CMD > gpurun_out/ncu_b.log 2>&1
python tools/curve_step.py > gpurun_out/${R}_curve_plain.log 2>&1 && ncu --set full --clock-control none --import-source on \
    -k regex:k_sim_paths -s 1 -c 1 -o gpurun_out/${R}_prof_curve_export -f python tools/curve_step.py > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/${R}_curve_plain.log
ls -la gpurun_out | tail -15
