#!/bin/bash
# symbolic hamEqs + resident-wave launches + PDL: parity, bench A/B, extras, rcp accuracy experiment
mkdir -p gpurun_out/r1c
O=gpurun_out/r1c
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 3500 $O/bench_n1.json; tail -3 $O/bench_n1.err
for cfg in "HB_NO_PDL=1" "HB_GRID_WAVES=2" "HB_GRID_WAVES=8" "HB_GRID_WAVES=1000" "HB_NO_PDL=1 HB_GRID_WAVES=1000"; do
  echo "== $cfg"; env $cfg python bench.py --steps 300 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j.get('fused16',{}).get('value'))"
done
python profiles/extras.py > $O/extras.log 2>&1; tail -20 $O/extras.log; cp gpurun_out/extras.json $O/ 2>/dev/null
profiles/exp/exp_rcp
