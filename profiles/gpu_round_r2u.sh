#!/bin/bash
# round 2, pass u: chain12 with the RK4 stages as one loop (instruction cache) against the unrolled build, FP64 operand-reuse microbenchmark, ncu digest
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 120 profiles/exp/exp_fp64b > $O/fp64_operands.txt 2>&1; cat $O/fp64_operands.txt
HB_AB_BUILTIN=1 timeout 900 python profiles/exp/exp_r2_ab.py sweep chain12 18 default,8af0be2 > $O/ab_chain12.txt 2>&1; cat $O/ab_chain12.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain12_step_rk4 -s 2 -c 1 -o $O/ncu_full_chain12 -f python profiles/ncu_target.py 1 4 7 18 > $O/ncu_chain12.log 2>&1; tail -1 $O/ncu_chain12.log
timeout 600 python -m pytest tests -m gpu -x -q -k "chain or full_size or n16 or 16" > $O/pytest_chain.log 2>&1; tail -3 $O/pytest_chain.log
