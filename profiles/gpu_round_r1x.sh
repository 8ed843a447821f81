#!/bin/bash
# full-size property tests of BASELINE configs 3-5, and the secondary configs once more (evolve numbers differed between boxes)
mkdir -p gpurun_out/r1x
O=gpurun_out/r1x
( time python -m pytest tests -m gpu -x -q -k "full_size" ) > $O/pytest_full_size.log 2>&1; tail -5 $O/pytest_full_size.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv | tee $O/smi.txt
python profiles/extras.py > $O/extras.log 2>&1; tail -20 $O/extras.log | cut -c1-220; cp gpurun_out/extras.json $O/ 2>/dev/null
