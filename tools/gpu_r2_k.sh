#!/bin/bash
# round-2 GPU call K (8 GPUs): bench.py --gpus 8 exactly as the driver launches it (ONE C3 problem row-sharded x8, c5 batch
# shards x8, the c4 block), and the reference arm's multi-rank behaviour (rank 0 only)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/k_gpus.txt
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/k_bench8.json 2> gpurun_out/k_bench8.err; echo "rc=$?" >> gpurun_out/k_bench8.err )
tail -5 gpurun_out/k_bench8.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/k_bench8.json").read().strip().splitlines()[-1])
    print("N=8 ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l.get("row_sharding"))
    print("c5", l.get("c5_batch")); print("c4", l.get("c4"))
except Exception as e:
    print("no line", e)
PY
