#!/bin/bash
# RKF45 controller shortcut (no divisions / pow when every error ratio is tiny): parity suite + reference-semantics throughput
mkdir -p gpurun_out/r1ab
O=gpurun_out/r1ab
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python profiles/extras.py > $O/extras.log 2>&1; grep -E "rkf45|evolveHam" $O/extras.log | cut -c1-200; cp gpurun_out/extras.json $O/ 2>/dev/null
