#!/bin/bash
# round 2, pass 3f: LDL^T without the 66 recomputed L d products — chain12 against the previous build; parity on the LDL^T systems
mkdir -p gpurun_out/r3f
O=gpurun_out/r3f
HB_AB_BUILTIN=1 HB_AB_LIBS="previous build (r3b)=profiles/ab_libs/lib_r3b.so" timeout 900 python profiles/exp/exp_r2_ab.py sweep chain12 18 default,previous > $O/ab_chain12.txt 2>&1; cat $O/ab_chain12.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "(chain or n16 or tape or singular or flags or rk4_step_vs or rkf45) and not full_size" > $O/pytest_subset.log 2>&1; tail -3 $O/pytest_subset.log
timeout 600 python -m pytest tests -m gpu -x -q -k "full_size_batches_against_the_oracle and chain" >> $O/pytest_subset.log 2>&1; tail -2 $O/pytest_subset.log
