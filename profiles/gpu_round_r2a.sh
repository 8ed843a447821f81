#!/bin/bash
# round 2, pass a: new tile-scheduled engine (2048-entry table by cp.async.bulk, 12-instruction sincos, ping-pong buffers,
# straight-line one-step instance): parity suite, FP64 issue-model microbenchmark, bench line, variant A/B, ncu capture
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
./profiles/exp/exp_fp64c > $O/fp64_issue_model.txt 2>&1; tail -3 $O/fp64_issue_model.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 1500 python profiles/exp/exp_r2_ab.py sweep double_pendulum 20 > $O/ab_double_pendulum.txt 2>&1; cat $O/ab_double_pendulum.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
