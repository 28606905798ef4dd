#!/bin/bash
# round-2 GPU call C: virtual ranks with the launch gate, sweep A/B incl. the 4-rows-per-warp kernel, device-resident HK
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "virtual or hopcroft" > gpurun_out/c_virtual_hk.log 2>&1; echo "rc=$?" >> gpurun_out/c_virtual_hk.log )
tail -6 gpurun_out/c_virtual_hk.log
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sweep or hopcroft" > gpurun_out/c_sweep_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c_sweep_tests.log )
tail -3 gpurun_out/c_sweep_tests.log
( timeout 300 python tools/gpu_sweep.py 30 > gpurun_out/c_sweep_ab.log 2>&1 )
cat gpurun_out/c_sweep_ab.log
( timeout 600 python tools/gpu_hk.py > gpurun_out/c_hk.log 2>&1 )
cat gpurun_out/c_hk.log
