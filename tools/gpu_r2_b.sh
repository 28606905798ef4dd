#!/bin/bash
# round-2 GPU call B: virtual-rank row sharding, sweep A/B, sanitizer runs on the cluster regime's failing soak seeds
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "virtual" > gpurun_out/b_virtual.log 2>&1; echo "rc=$?" >> gpurun_out/b_virtual.log )
tail -4 gpurun_out/b_virtual.log
( timeout 300 python tools/gpu_sweep.py 30 > gpurun_out/b_sweep_ab.log 2>&1 )
cat gpurun_out/b_sweep_ab.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep" > gpurun_out/b_sweep_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b_sweep_tests.log )
tail -3 gpurun_out/b_sweep_tests.log
export SSLAPB_CLUSTER_ANY_TSMALL=1
( timeout 500 compute-sanitizer --tool initcheck --print-limit 20 python tools/gpu_soak.py 14 4242 cluster_any > gpurun_out/b_san_initcheck.log 2>&1 )
tail -5 gpurun_out/b_san_initcheck.log
( timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_soak.py 14 4242 cluster_any > gpurun_out/b_san_memcheck.log 2>&1 )
tail -5 gpurun_out/b_san_memcheck.log
unset SSLAPB_CLUSTER_ANY_TSMALL
( timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize_case.py > gpurun_out/b_san_racecheck.log 2>&1 )
tail -5 gpurun_out/b_san_racecheck.log
