#!/bin/bash
# round 2, pass 3j: the committed build (after the like-term / closed-form changes) — full parity suite, smoke, bench line + reference arm
mkdir -p gpurun_out/r3j
O=gpurun_out/r3j
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; grep -E "passed|failed" $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3j/bench_n1.json'))
print('value %.4g  ms %.5f  frac %.3f  burst %.4g  e2e %.4g  fused16 %.4g  chain %.4g  hamEqs %.4g (%.3f)  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['burst']['value'], d['e2e']['value'], d['fused16']['value'], d['chain']['value'], d['ham_eqs']['value'], d['ham_eqs']['roofline']['frac'], d['clocks']))
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['configs'].items() if isinstance(v,dict)})
r=json.load(open('gpurun_out/r3j/bench_ref.json')); print('reference arm %.4g steps/s on %d threads' % (r['value'], r['cpu_baseline']['cores']))
PY
