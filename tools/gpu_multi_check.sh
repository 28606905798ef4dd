#!/bin/bash
# Multi-GPU check (gpurun --gpus 2): the 2-process NVLink test and bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/y_multi.log 2>&1; echo "rc=$?" >> gpurun_out/y_multi.log ); tail -3 gpurun_out/y_multi.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/y_bench2.json 2> gpurun_out/y_bench2.err; echo "rc=$?" >> gpurun_out/y_bench2.err ); tail -2 gpurun_out/y_bench2.err
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/y_ref2.json 2> gpurun_out/y_ref2.err; echo "rc=$?" >> gpurun_out/y_ref2.err ); tail -1 gpurun_out/y_ref2.err; cut -c1-300 gpurun_out/y_ref2.json
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/y_bench2.json").read().strip().splitlines()[-1])
    print("N=2 ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l.get("row_sharding"), (l.get("c5_batch") or {}).get("problems_per_s"), l["scaling"], l["config"]["parallelism"])
except Exception as e:
    print("no line", e)
PY
