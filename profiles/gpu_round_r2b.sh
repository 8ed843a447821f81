#!/bin/bash
# round 2, pass b: cp.async-staged Phase loads (dynamic shared memory layout) vs the r2a build (register ping-pong) on ONE box;
# CTA sizes for 6 warps/scheduler; parity suite; the new bench.py; ncu capture
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 1500 python profiles/exp/exp_r2_ab.py sweep double_pendulum 20 > $O/ab_double_pendulum.txt 2>&1; cat $O/ab_double_pendulum.txt
timeout 600 python profiles/exp/exp_r2_ab.py sweep triple_pendulum 20 "default,r2a build" > $O/ab_triple_pendulum.txt 2>&1; cat $O/ab_triple_pendulum.txt
timeout 900 python bench.py --steps 200 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
