#!/bin/bash
# A/B of two builds of the library on the same box: the in-tree one against abtest/*.so (SSLAP_B200_LIB), C3 regime profile
mkdir -p gpurun_out
for rep in 1 2; do
  for lib in abtest/*.so sslap_b200/csrc/libsslap_b200.so; do
    echo "== $lib" >> gpurun_out/ab.log
    ( SSLAP_B200_LIB=$PWD/$lib timeout 300 python tools/gpu_prof.py c3only >> gpurun_out/ab.log 2>&1 )
  done
done
grep -E "^==|^\[C3\]|per-round" gpurun_out/ab.log
