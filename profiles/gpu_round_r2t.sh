#!/bin/bash
# round 2, pass t (final build): every benchmark system against the round-1 build on one box, parity suite, smoke, bench line
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
for s in "double_pendulum 20" "triple_pendulum 20" "pendulum 21" "two_body 21" "spring1d 21"; do
  set -- $s
  timeout 600 python profiles/exp/exp_r2_ab.py sweep $1 $2 > $O/ab_$1.txt 2>&1; echo "== $1"; cut -c1-200 $O/ab_$1.txt
done
HB_AB_BUILTIN=1 timeout 600 python profiles/exp/exp_r2_ab.py sweep chain12 18 > $O/ab_chain12.txt 2>&1; echo "== chain12"; cut -c1-200 $O/ab_chain12.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; grep -E "passed|failed" $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2t/bench_n1.json'))
print('value %.4g  ms %.5f  frac %.3f  burst %.4g  e2e %.4g  fused16 %.4g  chain %.4g  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['burst']['value'], d['e2e']['value'], d['fused16']['value'], d['chain']['value'], d['clocks']))
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['configs'].items() if isinstance(v,dict)})
PY
