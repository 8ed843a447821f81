#!/bin/bash
# round 2, pass 3i: like-term collection in the system compiler + closed-form solves with the diagonal products folded —
# double / triple pendulum and chain12 against the previous build (r3h); quick parity subset
mkdir -p gpurun_out/r3i
O=gpurun_out/r3i
for s in "double_pendulum 20" "triple_pendulum 20" "chain12 18"; do
  set -- $s
  HB_AB_BUILTIN=1 HB_AB_LIBS="previous build (r3h)=profiles/ab_libs/lib_r3h.so" timeout 600 python profiles/exp/exp_r2_ab.py sweep $1 $2 default,previous > $O/ab_$1.txt 2>&1; echo "== $1"; cut -c1-230 $O/ab_$1.txt
done
timeout 600 python -m pytest tests -m gpu -x -q -k "rk4_step_vs or anchor or maps_and_energies or single_trajectory or exp_takes or out_of_domain" > $O/pytest_subset.log 2>&1; tail -2 $O/pytest_subset.log
