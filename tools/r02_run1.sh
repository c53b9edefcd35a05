#!/bin/bash
# Round-2 GPU session 1 (ONE B200, run under gpurun): GPU test suite, kernel variants, bench line, reference arm,
# ncu launch list + one full capture of the hot kernel.  Outputs under gpurun_out/.
set -x
R=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${R}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_tests.log
tail -5 gpurun_out/${R}_tests.log
for v in $(ls tools/variants/libmcts_*.so 2>/dev/null); do
  n=$(basename $v .so)
  MCTS_LIB=$v timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/${R}_var_${n}.log 2>&1
  echo "== $n"; tail -2 gpurun_out/${R}_var_${n}.log
done
timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/${R}_var_shipped.log 2>&1; echo "== shipped"; tail -2 gpurun_out/${R}_var_shipped.log
timeout 900 python bench.py > gpurun_out/${R}_bench_1gpu.json 2> gpurun_out/${R}_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
timeout 300 $CMD > gpurun_out/plain_a.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sim_paths -s 12 -c 1 \
    -o gpurun_out/${R}_prof_bench_sim_paths -f $CMD > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -20
head -c 3000 gpurun_out/${R}_bench_1gpu.json
