#!/bin/bash
# staged host pipeline (3 decoupled streams): parity + chunk-count sweep of the e2e leg; register-cap A/B of the RK4 kernel
mkdir -p gpurun_out/r1g
O=gpurun_out/r1g
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
for ch in 0 4 8 16 32 64; do
  echo -n "HB_HOST_CHUNKS=$ch : "; env HB_HOST_CHUNKS=$ch python bench.py --steps 300 --cpu-seconds 0.1 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4g  e2e %.4g steps/s' % (j['value'], j['e2e']['value']))"
done 2>&1 | tee $O/e2e_chunks.txt
for d in "" "HB_MINB_RK4=7" "HB_MINB_RK4=8"; do HB_JIT_DEFINES="$d" python profiles/exp/exp_jit_ab.py double_pendulum 2>&1 | tail -2; done | tee $O/minb_ab.txt
