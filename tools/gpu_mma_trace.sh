#!/bin/bash
tag=${1:-mt}
out=gpurun_out/$tag
mkdir -p $out
for lib in build/alt/libtrace_*.so; do
  name=$(basename $lib .so)
  sv=langevin; case $name in *dla*) sv=dl_adam;; *dl*) sv=dl;; esac
  CCVM_B200_LIB=$PWD/$lib timeout 120 python tools/mma_trace.py $sv > $out/$name.txt 2>&1
  echo "== $name"; sed -n 1,6p $out/$name.txt
done
