#!/bin/bash
# code-placement A/B: C3 with the mid regime off on the in-tree build and every build in abtest/ (per-round time of the few-bidder loops)
mkdir -p gpurun_out; rm -f gpurun_out/m_ab2.log
for lib in sslap_b200/csrc/libsslap_b200.so abtest/*.so; do echo "== $lib" >> gpurun_out/m_ab2.log; ( SSLAP_B200_LIB=$PWD/$lib timeout 100 python tools/gpu_prof.py c3mid0 >> gpurun_out/m_ab2.log 2>&1 ); done
grep -E "^==|^\[C|per-round" gpurun_out/m_ab2.log
