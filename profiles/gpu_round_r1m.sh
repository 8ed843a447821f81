#!/bin/bash
# A/B: sin r as fma(r, z*ps, r) (2 distinct register operands) vs fma(z*r, ps, r) (3), with and without the 80-register cap
mkdir -p gpurun_out/r1m
O=gpurun_out/r1m
for rep in 1 2; do
for d in "" "HB_MINB_RK4=1" "HB_SR_OLD=1" "HB_SR_OLD=1,HB_MINB_RK4=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done
done | tee $O/ab_sr.txt
for d in "" "HB_MINB_RK4=1" "HB_SR_OLD=1,HB_MINB_RK4=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py pendulum 21 2>&1 | tail -2; HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py two_body 21 2>&1 | tail -2; HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py triple_pendulum 2>&1 | tail -2; done | tee $O/ab_sr_others.txt
