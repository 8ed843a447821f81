#!/bin/bash
# round 2, pass 3c (final validation of this build): headline A/B against the previous build, parity suite, smoke, bench line +
# reference arm, ncu launch list of the bench command, ncu --set full capture of the headline kernel
mkdir -p gpurun_out/r3d
O=gpurun_out/r3d
HB_AB_BUILTIN=1 HB_AB_LIBS="previous build (r3b)=profiles/ab_libs/lib_r3b.so" timeout 600 python profiles/exp/exp_r2_ab.py sweep double_pendulum 20 "default,previous,round-1" > $O/ab_double_pendulum.txt 2>&1; cut -c1-220 $O/ab_double_pendulum.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3d/bench_n1.json'))
print('value %.4g  ms %.5f  frac %.3f  burst %.4g  e2e %.4g  fused16 %.4g  chain %.4g  hamEqs %.4g (%.3f)  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['burst']['value'], d['e2e']['value'], d['fused16']['value'], d['chain']['value'], d['ham_eqs']['value'], d['ham_eqs']['roofline']['frac'], d['clocks']))
print({k:(v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['configs'].items() if isinstance(v,dict)})
r=json.load(open('gpurun_out/r3d/bench_ref.json')); print('reference arm %.4g steps/s on %d threads' % (r['value'], r['cpu_baseline']['cores']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 20 --warmup 3 --no-extras --cpu-seconds 1 > $O/bench_under_ncu.json 2>> $O/bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_rk4 -s 3 -c 1 -o $O/ncu_full_step_rk4 -f python profiles/ncu_target.py 1 6 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
