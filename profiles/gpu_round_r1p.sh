#!/bin/bash
# A/B on one box: __launch_bounds__ of the RK4 step kernel -> registers 84 (none) / 120 (min 1 or 4) / 96 (min 5) / 80 (min 6)
mkdir -p gpurun_out/r1p
O=gpurun_out/r1p
for rep in 1 2; do
for d in "HB_MINB_RK4=0" "HB_MINB_RK4=1" "HB_MINB_RK4=5" "HB_MINB_RK4=6" "HB_MINB_RK4=0,HB_SR_OLD=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done
done | tee $O/ab_launch_bounds.txt
for s in triple_pendulum pendulum two_body; do l=20; [ $s != triple_pendulum ] && l=21
for d in "HB_MINB_RK4=0" "HB_MINB_RK4=1" "HB_MINB_RK4=3" "HB_MINB_RK4=6"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py $s $l 2>&1 | tail -2; done; done | tee $O/ab_launch_bounds_others.txt
