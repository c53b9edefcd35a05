#!/bin/bash
# Round-2 GPU session 3 (ONE B200): GPU tests on the rewritten replay kernel and the integer peak test, kernel variants.
set -x
R=r02
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${R}_tests3.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${R}_tests3.log
tail -8 gpurun_out/${R}_tests3.log
for v in $(ls tools/variants/libmcts_*.so 2>/dev/null); do
  n=$(basename $v .so)
  MCTS_PROFILE_REPLAY=227328 MCTS_LIB=$v timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/${R}_var3_${n}.log 2>&1
  echo "== $n"; tail -3 gpurun_out/${R}_var3_${n}.log
done
MCTS_PROFILE_REPLAY=227328 timeout 300 python tools/profile_step.py --reps 4 > gpurun_out/${R}_var3_shipped.log 2>&1; echo "== shipped"; tail -3 gpurun_out/${R}_var3_shipped.log
timeout 300 python tools/profile_step.py --reps 3 --dedup > gpurun_out/${R}_var3_dedup.log 2>&1; tail -1 gpurun_out/${R}_var3_dedup.log
