#!/bin/bash
# round 2, pass y: launch-shape sweep, remaining systems (reduced variant set; HB_SWEEP_FULL=1 for all 19)
mkdir -p gpurun_out/r2y
O=gpurun_out/r2y
for s in "two_body 21" "spring1d 21" "triple_pendulum 20" "room 20"; do
  set -- $s
  timeout 900 python profiles/exp/exp_shape_sweep.py sweep $1 $2 > $O/shape_$1.txt 2>&1; echo "== $1"; cat $O/shape_$1.txt
done
