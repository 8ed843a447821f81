#!/bin/bash
# consolidated round-1 pass of the final build (r1w: layout-specialised step kernel; r1r = the build before it): parity suite, smoke, both bench arms, the other configs, ncu launch list +
# full captures (double pendulum, triple pendulum, chain12), FP64/PCIe microbenchmarks
mkdir -p gpurun_out/r1w
O=gpurun_out/r1w
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 3800 $O/bench_n1.json; tail -3 $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 700 $O/bench_ref.json
python profiles/extras.py > $O/extras.log 2>&1; tail -20 $O/extras.log; cp gpurun_out/extras.json $O/ 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --cpu-seconds 0.2 > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4_fused16 -f python profiles/ncu_target.py 16 6 > $O/ncu_full16.log 2>&1; tail -1 $O/ncu_full16.log
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_triple -f python profiles/ncu_target.py 1 6 6 20 > $O/ncu_full_triple.log 2>&1; tail -1 $O/ncu_full_triple.log
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 2 -c 1 -o $O/ncu_full_chain12 -f python profiles/ncu_target.py 1 4 7 18 > $O/ncu_full_chain12.log 2>&1; tail -1 $O/ncu_full_chain12.log
ls -la $O
