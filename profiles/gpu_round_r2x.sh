#!/bin/bash
# round 2, pass x: launch shapes of the HBM-bound launches of systems that stage the 32 KB table image (every CTA stages its own)
mkdir -p gpurun_out/r2x
O=gpurun_out/r2x
for s in "double_pendulum 20" "pendulum 21"; do
  set -- $s
  timeout 900 python profiles/exp/exp_shape_sweep.py sweep $1 $2 > $O/shape_$1.txt 2>&1; echo "== $1"; cat $O/shape_$1.txt
done
