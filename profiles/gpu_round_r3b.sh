#!/bin/bash
# round 2, pass 3b: column-oriented (axpy) LDL^T + substitutions against the row-oriented build (saved), chain12; ncu digest; parity subset
mkdir -p gpurun_out/r3b
O=gpurun_out/r3b
HB_AB_BUILTIN=1 HB_AB_LIBS="row-oriented solve (saved build)=profiles/ab_libs/lib_r3_rowsolve.so" timeout 900 python profiles/exp/exp_r2_ab.py sweep chain12 18 default,row-oriented > $O/ab_chain12.txt 2>&1; cat $O/ab_chain12.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:chain12_step_rk4 -s 2 -c 1 -o $O/ncu_full_chain12 -f python profiles/ncu_target.py 1 4 7 18 > $O/ncu_chain12.log 2>&1; tail -1 $O/ncu_chain12.log
timeout 900 python -m pytest tests -m gpu -x -q -k "chain or n16 or full_size or tape or flags or singular" > $O/pytest_subset.log 2>&1; tail -3 $O/pytest_subset.log
