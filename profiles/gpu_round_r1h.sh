#!/bin/bash
# graph-cached host staging pipeline: parity test + chunk-count sweep with/without the graph, two batch sizes
mkdir -p gpurun_out/r1h
O=gpurun_out/r1h
( time python -m pytest tests -m gpu -x -q -k "host_chunk or layouts_and_memspaces or large_batch" ) > $O/pytest_host.log 2>&1; tail -5 $O/pytest_host.log
for g in 1 0; do for ch in 2 3 4 6 8 12 16 24 32; do HB_HOST_GRAPH=$g HB_HOST_CHUNKS=$ch python profiles/exp/exp_e2e.py 20 30 2>&1 | tail -1; done; done | tee $O/e2e_sweep.txt
for g in 1 0; do HB_HOST_GRAPH=$g python profiles/exp/exp_e2e.py 20 30 2>&1 | tail -1; HB_HOST_GRAPH=$g python profiles/exp/exp_e2e.py 22 10 2>&1 | tail -1;  HB_HOST_GRAPH=$g python profiles/exp/exp_e2e.py 17 100 2>&1 | tail -1; done | tee $O/e2e_auto.txt
python profiles/exp/exp_pcie.py 2>&1 | tail -8 | tee $O/pcie.txt
