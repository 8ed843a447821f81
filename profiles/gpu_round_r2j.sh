#!/bin/bash
# round 2, pass j: parity suite with the full-size oracle comparisons + evolveHamC + ensemble tests, bench line with the single-replay
# timing, reference arm
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/bench.err; tail -c 300 $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j/bench_n1.json'))
print('value %.4g  ms %.5f  frac %.3f  e2e %.4g  fused16 %.4g  chain %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['fused16']['value'], d['chain']['value']))
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['configs'].items() if isinstance(v,dict)})
print(d['roofline'].get('issue',{}).get('achieved_frac_of_issue_bound'), d['roofline'].get('issue',{}).get('achieved_frac_of_fp64_bound'))
PY
