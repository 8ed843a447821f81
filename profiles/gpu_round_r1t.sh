#!/bin/bash
# A/B: L2 look-ahead prefetch k iterations beyond the register prefetch (cold inputs: ring of buffers larger than L2)
mkdir -p gpurun_out/r1t
O=gpurun_out/r1t
for rep in 1 2; do
for d in "" "HB_L2_AHEAD=1" "HB_L2_AHEAD=2" "HB_L2_AHEAD=3"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2 | head -1; done
done | tee $O/ab_l2_ahead.txt
for s in pendulum two_body spring1d triple_pendulum; do l=21; [ $s = triple_pendulum ] && l=20
for d in "" "HB_L2_AHEAD=1" "HB_L2_AHEAD=2"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py $s $l 2>&1 | tail -2 | head -1; done; done | tee $O/ab_l2_ahead_others.txt
