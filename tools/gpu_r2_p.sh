#!/bin/bash
# round-2 GPU call P (1 GPU): full suite + bench.py on the current build
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/p_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/p_pytest.log )
tail -10 gpurun_out/p_pytest.log
( timeout 900 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "rc=$?" >> gpurun_out/p_bench.err ); tail -3 gpurun_out/p_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
    print("ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], "pageable", l["e2e_pageable"]["ms_per_step"])
    r = l["roofline"]; print("roofline", r["frac"], r["avg_launch_us"], "stream", r["streaming_only"]["frac"], r["streaming_only"]["avg_launch_us"])
    print(l["device_ms"]); print(l["c5_batch"]); print(l["cpu_baseline"])
except Exception as e:
    print("no line", e)
PY
( timeout 600 python tools/benchmarking.py --max-size 3200 --reference > gpurun_out/p_benchmarking.log 2>&1 ); tail -22 gpurun_out/p_benchmarking.log
