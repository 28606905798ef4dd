#!/bin/bash
# final confirmation on one B200: full -m gpu suite, smoke(), bench.py; then (GPU minutes permitting) the C3 regime profile and the
# C1-C3 timings through the drop-in call
mkdir -p gpurun_out; ( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/zf_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/zf_pytest.log ); tail -3 gpurun_out/zf_pytest.log; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-120; ( timeout 900 python bench.py > gpurun_out/zf_bench.json 2> gpurun_out/zf_bench.err; echo "rc=$?" >> gpurun_out/zf_bench.err ); tail -2 gpurun_out/zf_bench.err
( timeout 120 python tools/gpu_prof.py c3only > gpurun_out/zf_prof.log 2>&1 ); grep -A3 "^\[C3\]" gpurun_out/zf_prof.log | tail -4
( SKIP_C4=1 timeout 200 python tools/gpu_configs_time.py > gpurun_out/zf_configs.log 2>&1 ); cat gpurun_out/zf_configs.log
