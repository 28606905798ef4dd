#!/bin/bash
# round-2 GPU call D: HK device loop (fixed), virtual ranks K=2,4,8 with diagnostics, batch v2, stream probe
mkdir -p gpurun_out
( timeout 120 tools/probe/stream_probe > gpurun_out/d_probe.log 2>&1 ); cat gpurun_out/d_probe.log
( timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "virtual or hopcroft" > gpurun_out/d_virtual_hk.log 2>&1; echo "rc=$?" >> gpurun_out/d_virtual_hk.log )
grep -E "^E  .*Error|passed|failed|rc=" gpurun_out/d_virtual_hk.log | cut -c1-400 | tail -12
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hopcroft or batch or randomized_dense" > gpurun_out/d_hk_batch_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d_hk_batch_tests.log )
tail -4 gpurun_out/d_hk_batch_tests.log
( timeout 600 python tools/gpu_batch.py > gpurun_out/d_batch.log 2>&1 ); cat gpurun_out/d_batch.log
( timeout 600 python tools/gpu_hk.py > gpurun_out/d_hk.log 2>&1 ); cat gpurun_out/d_hk.log
