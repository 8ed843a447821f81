#!/bin/bash
# final validation of round 1: parity suite, bench line, one ncu --set full capture of the step kernel
mkdir -p gpurun_out/r1ad
O=gpurun_out/r1ad
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
