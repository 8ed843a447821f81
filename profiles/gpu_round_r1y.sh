#!/bin/bash
# A/B on the layout-specialised build: L2 look-ahead prefetch (cold inputs)
mkdir -p gpurun_out/r1y
O=gpurun_out/r1y
for rep in 1 2; do
for d in "" "HB_L2_AHEAD=1" "HB_L2_AHEAD=2"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2 | head -1; done
done | tee $O/ab_l2_ahead_layspec.txt
