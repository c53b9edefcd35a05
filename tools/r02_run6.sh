#!/bin/bash
# Round-2 GPU session 6 (TWO B200s): multi-device and NCCL-rank tests only.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s -k "group or device_list or nccl" > gpurun_out/r02_tests_multi_2gpu.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/r02_tests_multi_2gpu.log
tail -15 gpurun_out/r02_tests_multi_2gpu.log
