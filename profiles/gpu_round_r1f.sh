#!/bin/bash
# large-n path (dynamic smem, parked E/p, unrolled LDL^T) + zero-copy pinned-host path: parity, bench A/B, extras
mkdir -p gpurun_out/r1f
O=gpurun_out/r1f
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 2800 $O/bench_n1.json; tail -3 $O/bench_n1.err
echo "== HB_HOST_DIRECT=0"; HB_HOST_DIRECT=0 python bench.py --steps 300 --cpu-seconds 0.1 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['e2e'])"
python profiles/extras.py > $O/extras.log 2>&1; tail -20 $O/extras.log; cp gpurun_out/extras.json $O/ 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 2 -c 1 -o $O/ncu_full_chain12 -f python profiles/ncu_target.py 1 4 7 18 > $O/ncu_full_chain.log 2>&1; tail -2 $O/ncu_full_chain.log
