#!/bin/bash
# round 2, pass e: launch shapes for the memory-bound systems (pendulum, two-body, spring1d): the one-CTA-per-SM shape that wins
# for the double pendulum has too few bytes in flight for them (r2d: -24 % / -27 % against the round-1 build)
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 900 python profiles/exp/exp_r2_ab.py sweep pendulum 21 > $O/ab_pendulum.txt 2>&1; cat $O/ab_pendulum.txt
timeout 900 python profiles/exp/exp_r2_ab.py sweep two_body 21 > $O/ab_two_body.txt 2>&1; cat $O/ab_two_body.txt
