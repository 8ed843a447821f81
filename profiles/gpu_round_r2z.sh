#!/bin/bash
# round 2, pass z: the new launch shapes of the HBM-bound launches as the library's default (extras_r2 again), parity subset
mkdir -p gpurun_out/r2z
O=gpurun_out/r2z
timeout 900 python profiles/extras_r2.py $O > $O/extras_r2.log 2>&1; tail -32 $O/extras_r2.log | cut -c1-200
HB_AB_BUILTIN=1 timeout 600 python profiles/exp/exp_r2_ab.py sweep pendulum 21 default,8af0be2 > $O/ab_pendulum.txt 2>&1; cat $O/ab_pendulum.txt | cut -c1-250
timeout 900 python -m pytest tests -m gpu -x -q -k "maps_and_energies or ham_eqs or layouts or full_size or single_trajectory or cpp_mirror or flags" > $O/pytest_subset.log 2>&1; tail -3 $O/pytest_subset.log
