#!/bin/bash
# round 2, pass v: RKF45 stage loop of large systems (chain12 stepHam) against the unrolled build; one-evaluation kernels vs the HBM roofline; evolve paths
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
HB_AB_BUILTIN=1 timeout 900 python profiles/exp/exp_r2_ab.py sweep chain12 18 default,8af0be2 > $O/ab_chain12.txt 2>&1; cat $O/ab_chain12.txt
timeout 900 python profiles/extras_r2.py $O > $O/extras_r2.log 2>&1; tail -32 $O/extras_r2.log | cut -c1-330
timeout 900 python -m pytest tests -m gpu -x -q -k "chain or evolve or n16 or full_size" > $O/pytest_chain.log 2>&1; tail -3 $O/pytest_chain.log
