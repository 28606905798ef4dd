#!/bin/bash
# parity suites on the in-tree build, then A/B of the C3 regime profile against the builds dropped into abtest/
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "not c4 and not c5_full" > gpurun_out/parity.log 2>&1; echo "rc=$?" >> gpurun_out/parity.log )
tail -4 gpurun_out/parity.log
rm -f gpurun_out/ab.log; bash tools/gpu_ab.sh
