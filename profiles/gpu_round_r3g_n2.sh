#!/bin/bash
# round 2, pass 3g (2 GPUs): bench.py under torchrun at N = 2 (both arms), ensemble tests on 2 GPUs
mkdir -p gpurun_out/r3g
O=gpurun_out/r3g
timeout 600 python -m pytest tests -m gpu -x -q -k "ensemble" > $O/pytest_ensemble_n2.log 2>&1; tail -2 $O/pytest_ensemble_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1000 --warmup 10 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 600 $O/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/bench_ref_n2.json 2>> $O/bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3g/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.4g ms %.5f frac %.3f e2e %.4g hamEqs %.4g gather_ms %.3f with_gather %.4g c4 %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['ham_eqs']['value'], d.get('gather_ms', -1), d.get('value_with_gather', -1), {k:(v.get('value'), v.get('gather_ms')) for k,v in d['configs'].items() if isinstance(v, dict)}))
r=json.loads(open('gpurun_out/r3g/bench_ref_n2.json').read().strip().splitlines()[-1]); print('ref', r.get('value'), r.get('n_gpus'))
PY
