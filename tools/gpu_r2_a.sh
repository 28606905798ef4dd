#!/bin/bash
# round-2 GPU call A: full -m gpu suite, default bench, soak of the cluster regime incl. the combinations that failed in round 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/a_gpu.txt
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log ) 
tail -5 gpurun_out/a_pytest.log
( timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err )
tail -c 1500 gpurun_out/a_bench.json
( timeout 400 python tools/gpu_soak.py 500 777 plain > gpurun_out/a_soak_plain.log 2>&1 )
tail -2 gpurun_out/a_soak_plain.log
( SSLAPB_CLUSTER_ANY_TSMALL=1 timeout 500 python tools/gpu_soak.py 600 4242 cluster_any > gpurun_out/a_soak_cluster_any.log 2>&1 )
tail -3 gpurun_out/a_soak_cluster_any.log
