#!/bin/bash
# iteration traces of both warpgroups (DL, DL-adam; n = 70) with stagger / pacing off and on
tag=${1:-m21}
out=gpurun_out/$tag
mkdir -p $out
for mode in ${MODES:-00 10 11}; do
  s=${mode:0:1}; p=${mode:1:1}
  CCVM_MMA_STAGGER=$s CCVM_MMA_PACE=$p CCVM_B200_LIB=$PWD/build/alt/libtracedl.so timeout 120 python tools/mma_trace.py dl 70 > $out/trace_dl_m$mode.txt 2>$out/err.txt; echo "dl $mode rc=$?"
  sed -n 8,14p $out/trace_dl_m$mode.txt
  CCVM_MMA_STAGGER=$s CCVM_MMA_PACE=$p CCVM_B200_LIB=$PWD/build/alt/libtracedla.so timeout 120 python tools/mma_trace.py dl_adam 70 > $out/trace_dla_m$mode.txt 2>>$out/err.txt; echo "dla $mode rc=$?"
  sed -n 8,14p $out/trace_dla_m$mode.txt
done
tail -3 $out/err.txt
