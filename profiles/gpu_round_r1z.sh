#!/bin/bash
# A/B: register caps of the adaptive (reference-semantics) kernel: 214 registers (2 CTAs/SM) vs 168 / 128 / 96 with spills
mkdir -p gpurun_out/r1z
O=gpurun_out/r1z
for s in double_pendulum triple_pendulum pendulum; do
for d in "" "HB_MINB_RKF45=3" "HB_MINB_RKF45=4" "HB_MINB_RKF45=5"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_rkf45.py $s 2>&1 | tail -1; done; done | tee $O/ab_rkf45_regs.txt
