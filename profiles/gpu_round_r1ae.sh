#!/bin/bash
# last validation of round 1 (table staging without the trip-count division): parity suite + bench line
mkdir -p gpurun_out/r1ae
O=gpurun_out/r1ae
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python bench.py --cpu-seconds 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 300 $O/bench_n1.json; tail -3 $O/bench_n1.err
