#!/bin/bash
# round 2, pass c: CTA size picked by occupancy (384 / 256 / 128), load-path variants at CTA 384, e2e per variant (the r2b e2e
# regression), round-1 final build on the same box as the reference
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 1800 python profiles/exp/exp_r2_ab.py sweep double_pendulum 20 > $O/ab_double_pendulum.txt 2>&1; cat $O/ab_double_pendulum.txt
timeout 600 python profiles/exp/exp_r2_ab.py sweep triple_pendulum 20 "default,round-1,CTA 128,CTA 256" > $O/ab_triple_pendulum.txt 2>&1; cat $O/ab_triple_pendulum.txt
