#!/bin/bash
# profiling builds of the mid regime (abtest/lib_prof<k>.so, -DSSLAPB_MID_PROF=k: prof_ms[6] = 1 bidding / 2 merge / 3 compaction
# part of the rounds, 4 undecided bids (x 1e-3) + bids (x 1e-6)): C3 at the default t_mid.  The builds are made here, on the CPU box:
#   for k in 1 2 3 4; do d=build/tmp/p$k; mkdir -p $d/sslap_b200 $d/include; cp -r sslap_b200/csrc $d/sslap_b200/; cp include/sslap_b200.h $d/include/
#     make -C $d/sslap_b200/csrc clean libsslap_b200.so NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
#       -Xcompiler -fPIC --fmad=false -DSSLAPB_MID_PROF=$k" && cp $d/sslap_b200/csrc/libsslap_b200.so abtest/lib_prof$k.so; done
mkdir -p gpurun_out; rm -f gpurun_out/m_prof.log
for lib in abtest/lib_prof*.so; do echo "== $lib" >> gpurun_out/m_prof.log; ( SSLAP_B200_LIB=$PWD/$lib timeout 100 python tools/gpu_prof.py c3only >> gpurun_out/m_prof.log 2>&1 ); done
grep -E "^==|mid=" gpurun_out/m_prof.log
