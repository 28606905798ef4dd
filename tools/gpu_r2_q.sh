#!/bin/bash
# parity on the current build (tie-sensitive goldens, randomized differential, hot-list test, virtual ranks), then A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "not c4 and not c5_full" > gpurun_out/q_parity.log 2>&1; echo "rc=$?" >> gpurun_out/q_parity.log )
tail -5 gpurun_out/q_parity.log
rm -f gpurun_out/ab.log; bash tools/gpu_ab.sh
