#!/bin/bash
# zero-copy host path (warp-transposed stores): full GPU parity suite, e2e A/B (zero-copy vs staged), bench line
mkdir -p gpurun_out/r1i
O=gpurun_out/r1i
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
for n in 20 22 17; do
  python profiles/exp/exp_e2e.py $n 30 2>&1 | tail -1
  HB_HOST_DIRECT=0 python profiles/exp/exp_e2e.py $n 30 2>&1 | tail -1
done | tee $O/e2e_ab.txt
python bench.py --steps 1000 --cpu-seconds 5 > $O/bench_n1.json 2> $O/bench_n1.err; cat $O/bench_n1.json | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  e2e %.4g steps/s frac %.3f' % (j['value'], j['e2e']['value'], j['roofline']['frac']))"
