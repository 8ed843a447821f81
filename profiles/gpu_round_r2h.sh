#!/bin/bash
# round 2, pass h: does NVML polling perturb the timed launches? (bench at 0.5 / 10 / 50 ms polling), full parity suite after the
# kernel-cache fix, multi-config bench line
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
for ms in 0.5 10 50; do
  HB_BENCH_NVML_MS=$ms timeout 300 python bench.py --no-extras --cpu-seconds 1 > $O/bench_nvml_$ms.json 2>> $O/bench.err
  python -c "import json;d=json.load(open('$O/bench_nvml_$ms.json'));print('NVML poll $ms ms: %.3f us/launch, %.4g steps/s, samples %d' % (d['ms_per_step']*1e3, d['value'], d['clocks']['samples']))" | tee -a $O/nvml_polling.txt
done
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1.json 2>> $O/bench.err; tail -c 300 $O/bench_n1.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err; tail -c 300 $O/bench_ref.json
