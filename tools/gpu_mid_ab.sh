#!/bin/bash
# mid regime (auction.cu mid_regime): the t_mid sweep (0 = off, 64, 128, 256) on C3 (PROF_C2=1: and C2) for the in-tree build and
# every build dropped into abtest/ (SSLAP_B200_LIB) — its and a checksum of sol are printed per run; PARITY=1: parity subset first
mkdir -p gpurun_out; rm -f gpurun_out/m_ab.log
prof() { echo "== $1" >> gpurun_out/m_ab.log; ( SSLAP_B200_LIB=$PWD/$1 timeout 150 python tools/gpu_prof.py midsweep >> gpurun_out/m_ab.log 2>&1 ); }
if [ -n "$PARITY" ]; then
  ( timeout 300 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "mid_regime or hot_lists or golden or randomized_differential or c2_matches or very_long or max_iter" > gpurun_out/m_parity.log 2>&1; echo "rc=$?" >> gpurun_out/m_parity.log )
  tail -3 gpurun_out/m_parity.log
fi
prof sslap_b200/csrc/libsslap_b200.so
for lib in abtest/*.so; do prof $lib; done
grep -E "^==|^\[C|per-round|not settable" gpurun_out/m_ab.log
