#!/bin/bash
# host paths: parity of all four modes, then e2e of zero-copy vs hybrid (copy-engine upload + direct stores) vs staged
mkdir -p gpurun_out/r1n
O=gpurun_out/r1n
( time python -m pytest tests -m gpu -x -q -k "host_paths" ) > $O/pytest_host.log 2>&1; tail -5 $O/pytest_host.log
( python profiles/exp/exp_e2e.py 20 30 2>&1 | tail -1
for ch in 4 8 12 16 24 32; do HB_HOST_DIRECT=2 HB_HOST_CHUNKS=$ch python profiles/exp/exp_e2e.py 20 30 2>&1 | tail -1; done
HB_HOST_DIRECT=0 python profiles/exp/exp_e2e.py 20 30 2>&1 | tail -1
for n in 17 22; do python profiles/exp/exp_e2e.py $n 20 2>&1 | tail -1; HB_HOST_DIRECT=2 python profiles/exp/exp_e2e.py $n 20 2>&1 | tail -1; done ) | tee $O/e2e_modes.txt
