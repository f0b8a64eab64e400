#!/bin/bash
# decompose the xprop2 kernel time: full / no MMA / no loads / no epilogue / pure loads / pure MMA issue / no W / no X
mkdir -p gpurun_out
for v in ${VARIANTS:-2 3}; do
for a in 0 1 2 4 5 6 8 16 7; do
  echo "== variant $v BSMM_ABLATE=$a" | tee -a gpurun_out/ablate.txt
  BSMM_XPROP2=$v BSMM_ABLATE=$a timeout 120 python tools/time_ops.py ${DENS:-0.05 0.25} 2>&1 | sed 's/| bprop.*//' | tee -a gpurun_out/ablate.txt
done
done
