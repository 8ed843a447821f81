#!/bin/bash
# round 2, pass 3a: chain12 — number of waves (dynamic CTA scheduling) for the one-evaluation kernels and the step kernel
mkdir -p gpurun_out/r3a
O=gpurun_out/r3a
HB_SWEEP_K=50 HB_SWEEP_WAVES=0.5,1,2,4 timeout 900 python profiles/exp/exp_shape_sweep.py sweep chain12 18 > $O/shape_chain12.txt 2>&1; cat $O/shape_chain12.txt
