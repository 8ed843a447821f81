#!/bin/bash
# round 2, pass f: staging and launch shape by compute intensity (Sys::HEAVY): every benchmark system against the round-1 final
# build on one box; parity suite incl. the C-ABI ensembles; bench line; ncu capture of the headline kernel
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
for s in "double_pendulum 20" "triple_pendulum 20" "pendulum 21" "two_body 21" "spring1d 21" "chain12 18"; do
  set -- $s
  timeout 600 python profiles/exp/exp_r2_ab.py sweep $1 $2 > $O/ab_$1.txt 2>&1; echo "== $1"; cat $O/ab_$1.txt
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
