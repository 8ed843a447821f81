#!/bin/bash
# round 2, pass 3e: config 3 on two streams vs alternating on one stream
mkdir -p gpurun_out/r3e
timeout 600 python bench.py --config 3 --steps 200 > gpurun_out/r3e/bench_config3.json 2> gpurun_out/r3e/err.txt; python -c "
import json; d=json.load(open('gpurun_out/r3e/bench_config3.json'))['detail']; print(d['mode'], d['ms_per_step_two_streams'], d['ms_per_step_one_stream'], d['value'], d['roofline']['frac'])"
timeout 600 python -m pytest tests -m gpu -x -q -k "exp_takes" 2>&1 | tail -2
