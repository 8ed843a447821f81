#!/bin/bash
# validation of the final build: parity suite, smoke, bench, secondary configs
mkdir -p gpurun_out/r1aa
O=gpurun_out/r1aa
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 600 $O/bench_n1.json; tail -3 $O/bench_n1.err
python profiles/extras.py > $O/extras.log 2>&1; cut -c1-200 $O/extras.log; cp gpurun_out/extras.json $O/ 2>/dev/null
