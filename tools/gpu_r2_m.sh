#!/bin/bash
# round-2 GPU call M: lean tail loops — parity (hot-list test, randomized differential, goldens), then A/B against other builds
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "not c4 and not c5_full" > gpurun_out/m_parity.log 2>&1; echo "rc=$?" >> gpurun_out/m_parity.log )
tail -6 gpurun_out/m_parity.log
( SSLAP_B200_LIB=$PWD/sslap_b200/csrc/libsslap_b200.so timeout 200 python tools/gpu_bisect.py 2>&1 | grep "hot=1" )
rm -f gpurun_out/ab.log; bash tools/gpu_ab.sh
