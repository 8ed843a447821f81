#!/bin/bash
# round 2, pass n: the light (HBM-bound) systems and the chain against the round-1 build: register prefetch on the plain-load path,
# one vs two resident waves
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
for s in "spring1d 21" "pendulum 21" "two_body 21"; do
  set -- $s
  timeout 600 python profiles/exp/exp_r2_ab.py sweep $1 $2 > $O/ab_$1.txt 2>&1; echo "== $1"; cat $O/ab_$1.txt
done
HB_AB_BUILTIN=1 timeout 600 python profiles/exp/exp_r2_ab.py sweep chain12 18 "default (register,2 resident,round-1" > $O/ab_chain12.txt 2>&1; echo "== chain12"; cat $O/ab_chain12.txt
