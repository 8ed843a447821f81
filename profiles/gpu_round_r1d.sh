#!/bin/bash
# launch-shape sweep (CTA size x resident waves) for the config-2 bench; SPD3 A/B; parity
mkdir -p gpurun_out/r1d
O=gpurun_out/r1d
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for blk in 128 160 192 256 384; do for w in 1 1.5 2 3 4; do
  echo -n "block=$blk waves=$w : "; env HB_BLOCK=$blk HB_GRID_WAVES=$w python bench.py --steps 300 --cpu-seconds 0.1 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g steps/s  %.3f us/step  fused16 %.4g' % (j['value'], j['ms_per_step']*1e3, j.get('fused16',{}).get('value')))"
done; done 2>&1 | tee $O/sweep.txt
python profiles/exp/exp_jit_ab.py triple_pendulum 2>&1 | tail -2
HB_JIT_DEFINES="HB_SPD3_LDLT=1" python profiles/exp/exp_jit_ab.py triple_pendulum 2>&1 | tail -2
