#!/bin/bash
# round-2 GPU call G (1 GPU): hot lists — parity suites, then per-regime profile with the hot lists on / off
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/g_parity.log 2>&1; echo "rc=$?" >> gpurun_out/g_parity.log )
tail -15 gpurun_out/g_parity.log
( timeout 600 python tools/gpu_prof.py both > gpurun_out/g_prof.log 2>&1 ); cat gpurun_out/g_prof.log
( timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "not c4" > gpurun_out/g_configs.log 2>&1; echo "rc=$?" >> gpurun_out/g_configs.log )
tail -15 gpurun_out/g_configs.log
