#!/bin/bash
# round 2, pass d: launch shape by the cost model (one CTA per SM, warps chosen for whole rounds), e2e back on plain loads,
# round-1 final and r2a builds on the same box as references; multi-config bench
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 1800 python profiles/exp/exp_r2_ab.py sweep double_pendulum 20 > $O/ab_double_pendulum.txt 2>&1; cat $O/ab_double_pendulum.txt
timeout 600 python profiles/exp/exp_r2_ab.py sweep triple_pendulum 20 "default,round-1,CTA 512,CTA 384" > $O/ab_triple_pendulum.txt 2>&1; cat $O/ab_triple_pendulum.txt
timeout 600 python profiles/exp/exp_r2_ab.py sweep pendulum 21 "default,round-1" > $O/ab_pendulum.txt 2>&1; cat $O/ab_pendulum.txt
timeout 600 python profiles/exp/exp_r2_ab.py sweep two_body 21 "default,round-1" > $O/ab_two_body.txt 2>&1; cat $O/ab_two_body.txt
