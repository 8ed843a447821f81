#!/bin/bash
# A/B: sincos constants as literals / immediate magic constant (fewer three-register DFMAs: 48 -> 32 per step statically)
mkdir -p gpurun_out/r1ac
O=gpurun_out/r1ac
for rep in 1 2; do
for d in "" "HB_SC_MAGIC_IMM=1" "HB_SC_MAGIC_IMM=1,HB_SC_LITERALS=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done
done | tee $O/ab_sc_literals.txt
for s in pendulum triple_pendulum; do l=21; [ $s = triple_pendulum ] && l=20
for d in "" "HB_SC_MAGIC_IMM=1,HB_SC_LITERALS=1"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py $s $l 2>&1 | tail -2; done; done | tee $O/ab_sc_literals_others.txt
