#!/bin/bash
# round-2 GPU call E (2 GPUs): the real NVLink path — 2-process row-sharded solve test, virtual ranks K=4,8 after the gate fix,
# bench.py --gpus 2 (strong scaling of ONE C3 problem) at two t_shard settings
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_gpus.txt 2>&1; nvidia-smi topo -m >> gpurun_out/e_gpus.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/e_multi.log 2>&1; echo "rc=$?" >> gpurun_out/e_multi.log )
tail -5 gpurun_out/e_multi.log
( timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -q -k "virtual" > gpurun_out/e_virtual.log 2>&1; echo "rc=$?" >> gpurun_out/e_virtual.log )
grep -E "^E  .*rank|passed|failed|rc=" gpurun_out/e_virtual.log | cut -c1-300 | tail -12
for ts in 16384 2048; do
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --t-shard $ts > gpurun_out/e_bench2_$ts.json 2> gpurun_out/e_bench2_$ts.err; echo "rc=$?" >> gpurun_out/e_bench2_$ts.err )
tail -3 gpurun_out/e_bench2_$ts.err; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/e_bench2_$ts.json").read().strip().splitlines()[-1])
    print("N=2 t_shard=$ts ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], l.get("row_sharding"), (l.get("c5_batch") or {}).get("problems_per_s"))
except Exception as e:
    print("no line", e)
PY
done
