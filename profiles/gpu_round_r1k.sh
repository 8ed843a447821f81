#!/bin/bash
# A/B (NVRTC builds of the same engine, same process conditions): transposed-store code present or not, register prefetch vs
# L2 prefetch of the next Phase, register caps -> occupancy 6/7/8 CTAs per SM (double pendulum), 4/5 (triple pendulum)
mkdir -p gpurun_out/r1k
O=gpurun_out/r1k
for rep in 1 2; do
for d in "" "HB_WSTORE_MAXD=0" "HB_L2_PREFETCH=1" "HB_L2_PREFETCH=1,HB_MINB_RK4=7" "HB_L2_PREFETCH=1,HB_MINB_RK4=8"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done
done | tee $O/ab_double_pendulum.txt
for d in "" "HB_L2_PREFETCH=1" "HB_L2_PREFETCH=1,HB_MINB_RK4=5" "HB_MINB_RK4=5"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py triple_pendulum 2>&1 | tail -2; done | tee $O/ab_triple_pendulum.txt
for d in "" "HB_L2_PREFETCH=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py pendulum 21 2>&1 | tail -2; HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py two_body 21 2>&1 | tail -2; done | tee $O/ab_small.txt
