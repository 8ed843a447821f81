#!/bin/bash
# round 2, pass w: table-driven exp (hb_exp) — room / spring / bezier against the build with libdevice's exp; headline kernel re-checked (table image grew by 512 B)
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
for s in "room 20" "spring 20" "bezier 21" "double_pendulum 20"; do
  set -- $s
  HB_AB_BUILTIN=1 timeout 600 python profiles/exp/exp_r2_ab.py sweep $1 $2 default,8af0be2 > $O/ab_$1.txt 2>&1; echo "== $1"; cut -c1-260 $O/ab_$1.txt
done
timeout 900 python -m pytest tests -m gpu -x -q -k "room or spring or bezier or tape or flags or domain or layouts or rk4_step_vs or rkf45" > $O/pytest_subset.log 2>&1; tail -3 $O/pytest_subset.log
