#!/bin/bash
# A/B: step_rk4 dispatching once per launch into a layout-specialised body (no SOA code on the AOS path) vs one generic body
mkdir -p gpurun_out/r1u
O=gpurun_out/r1u
for rep in 1 2; do
for d in "HB_LAYSPEC=0" "HB_LAYSPEC=1" "HB_LAYSPEC=1,HB_MINB_RK4=6"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done
done | tee $O/ab_layspec.txt
for s in pendulum two_body spring1d triple_pendulum; do l=21; [ $s = triple_pendulum ] && l=20
for d in "HB_LAYSPEC=0" "HB_LAYSPEC=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py $s $l 2>&1 | tail -2 | head -1; done; done | tee $O/ab_layspec_others.txt
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
