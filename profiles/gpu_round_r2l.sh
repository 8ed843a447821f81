#!/bin/bash
# round 2, pass l: chain12 LDL^T ordering A/B (saved builds, one box), ncu digests of step_rkf45 and of the chain12 step kernel
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 900 python profiles/exp/exp_r2_ab.py sweep chain12 18 > $O/ab_chain12_ldlt.txt 2>&1; cat $O/ab_chain12_ldlt.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_rkf45 -s 2 -c 1 -o $O/ncu_full_step_rkf45 -f python profiles/ncu_target_rkf45.py 4 > $O/ncu_rkf45.log 2>&1; tail -1 $O/ncu_rkf45.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain12_step_rk4 -s 2 -c 1 -o $O/ncu_full_chain12 -f python profiles/ncu_target.py 1 4 7 18 > $O/ncu_chain12.log 2>&1; tail -1 $O/ncu_chain12.log
