#!/bin/bash
# A/B: sin/cos table staged 1x / 2x / 4x (lane l reads copy l % R: fewer shared-memory bank conflicts per LDS.128)
mkdir -p gpurun_out/r1l
O=gpurun_out/r1l
for s in double_pendulum pendulum triple_pendulum; do
for d in "" "HB_SC_REP_LOG2=1" "HB_SC_REP_LOG2=2"; do l=20; [ $s = pendulum ] && l=21; HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py $s $l 2>&1 | tail -2; done
done | tee $O/ab_screp.txt
