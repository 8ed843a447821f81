#!/bin/bash
# symbolic hamEqs + 2 resident waves + PDL + cubic rcp: parity, both bench arms, extras, ncu launch list + full captures, PCIe
mkdir -p gpurun_out/r1e
O=gpurun_out/r1e
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 3500 $O/bench_n1.json; tail -3 $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
python profiles/extras.py > $O/extras.log 2>&1; tail -20 $O/extras.log; cp gpurun_out/extras.json $O/ 2>/dev/null
python profiles/exp/exp_pcie.py 2>&1 | tee $O/pcie.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 > $O/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_triple -f python profiles/ncu_target.py 1 6 6 20 > $O/ncu_full_triple.log 2>&1; tail -1 $O/ncu_full_triple.log
ls -la $O
