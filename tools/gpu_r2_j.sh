#!/bin/bash
# round-2 GPU call J (1 GPU): bench.py with the hot form, ncu launch list of bench.py, ncu --set full of the sweep kernels,
# small-problem timings with the grid heuristic
mkdir -p gpurun_out
( timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "rc=$?" >> gpurun_out/j_bench.err ); tail -3 gpurun_out/j_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/j_bench.json").read().strip().splitlines()[-1])
    print("ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], "pageable", l["e2e_pageable"]["ms_per_step"])
    r = l["roofline"]; print("roofline", r["frac"], r["avg_launch_us"], "stream", r["streaming_only"]["frac"], r["streaming_only"]["avg_launch_us"], "insitu", r["insitu"])
    print(l["device_ms"]); print(l["c5_batch"]); print(l["cpu_baseline"])
except Exception as e:
    print("no line", e)
PY
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_bench.log 2>&1 ); tail -2 gpurun_out/j_ncu_bench.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:bid_sweep -c 16 -o gpurun_out/j_sweep_full python tools/gpu_sweep.py 1 > gpurun_out/j_ncu_sweep.log 2>&1 ); tail -3 gpurun_out/j_ncu_sweep.log
( timeout 300 python tools/gpu_small.py > gpurun_out/j_small.log 2>&1 ); grep "max_ctas=  0" gpurun_out/j_small.log
