#!/bin/bash
# launch shape of the cheap systems: grid = waves x resident CTAs, with and without the layout-specialised body
mkdir -p gpurun_out/r1v
O=gpurun_out/r1v
for s in pendulum two_body spring1d; do
for ls in 0 1; do for w in 1 1.5 2 3 4; do echo -n "waves=$w "; HB_GRID_WAVES=$w HB_JIT_DEFINES="HB_LAYSPEC=$ls" python profiles/exp/exp_jit_ab.py $s 21 2>&1 | tail -2 | head -1; done; done
done | tee $O/waves_small_systems.txt
