#!/bin/bash
# round-2 GPU call F (1 GPU): the full -m gpu suite on the current binary, per-regime profile C2 vs C3 (is the tail's row load a
# DRAM miss?), bench.py, ncu launch list of bench.py, ncu --set full of the sweep kernel variants
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log )
tail -14 gpurun_out/f_pytest.log
( timeout 300 python tools/gpu_prof.py both > gpurun_out/f_prof.log 2>&1 ); cat gpurun_out/f_prof.log
( timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?" >> gpurun_out/f_bench.err ); tail -2 gpurun_out/f_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/f_bench.json").read().strip().splitlines()[-1])
    print("ms/step", l["ms_per_step"], "e2e", l["e2e"]["ms_per_step"], "roofline", l["roofline"]["frac"], l["roofline"]["avg_launch_us"], "c5", l["c5_batch"]["batch_kernel_ms"], l["c5_batch"]["problems_per_s"])
    print(l["device_ms"]); print(l["cpu_baseline"])
except Exception as e:
    print("no line", e)
PY
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1 ); tail -2 gpurun_out/f_ncu_bench.log
( timeout 600 ncu --set full --clock-control none --import-source on -k regex:bid_sweep -c 6 -o gpurun_out/f_sweep_full python tools/gpu_sweep.py 1 > gpurun_out/f_ncu_sweep.log 2>&1 ); tail -3 gpurun_out/f_ncu_sweep.log
ls -la gpurun_out/*.ncu-rep
