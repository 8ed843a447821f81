#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
python profiles/exp_minb.py 2>&1 | grep "dp jit"
python profiles/exp_block.py all 2>&1 | tail -12
