#!/bin/bash
# round-2 GPU call H (1 GPU): hot lists in every regime + long-row instance + stand-alone hot sweep: full suite, profile,
# sweep A/B, the reference-style benchmarking sweeps (sizes 10..3162, densities) against the reference
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log )
tail -12 gpurun_out/h_pytest.log
( timeout 600 python tools/gpu_prof.py c3 > gpurun_out/h_prof.log 2>&1 ); cat gpurun_out/h_prof.log
( timeout 300 python tools/gpu_sweep.py 30 > gpurun_out/h_sweep_ab.log 2>&1 ); cat gpurun_out/h_sweep_ab.log
( timeout 900 python tools/benchmarking.py --max-size 3200 --reference > gpurun_out/h_benchmarking.log 2>&1 ); cat gpurun_out/h_benchmarking.log
