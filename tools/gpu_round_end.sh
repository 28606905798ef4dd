#!/bin/bash
# What produces the committed round-end evidence on one B200 (profiles/r2_*): the full -m gpu suite, bench.py, the ncu launch
# list of bench.py, the ncu --set full captures of the sweep kernels, the regime profile, the config timings.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/z_pytest.log )
tail -9 gpurun_out/z_pytest.log
( timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "rc=$?" >> gpurun_out/z_bench.err ); tail -2 gpurun_out/z_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/z_bench.json").read().strip().splitlines()[-1])
    print("ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], "pageable", l["e2e_pageable"]["ms_per_step"])
    r = l["roofline"]; print("roofline", r["frac"], r["avg_launch_us"], "clean", r["clean_l2_flush"]["frac"], "stream", r["streaming_only"]["frac"], r["streaming_only"]["avg_launch_us"])
    print(l["device_ms"]); print(l["c5_batch"]); print(l["cpu_baseline"])
except Exception as e:
    print("no line", e)
PY
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/z_ncu_bench.log 2>&1 ); tail -1 gpurun_out/z_ncu_bench.log | cut -c1-200
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:bid_sweep -c 24 -o gpurun_out/z_sweep_full python tools/gpu_sweep.py 1 > gpurun_out/z_ncu_sweep.log 2>&1 ); tail -2 gpurun_out/z_ncu_sweep.log
( timeout 300 python tools/gpu_prof.py c3 > gpurun_out/z_prof.log 2>&1 ); grep -A3 "^\[C3\]" gpurun_out/z_prof.log | head -8
( timeout 600 python tools/gpu_configs_time.py > gpurun_out/z_configs.log 2>&1 ); cat gpurun_out/z_configs.log
