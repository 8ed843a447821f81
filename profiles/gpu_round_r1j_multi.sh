#!/bin/bash
# multi-GPU: the bench line at N = 2, 4, 8 (torchrun, NCCL) and BASELINE configs[3] (8M triple pendulums, gather at end)
mkdir -p gpurun_out/r1j
O=gpurun_out/r1j
G=$(nvidia-smi -L | wc -l); echo "GPUs: $G"
for n in 2 4 8; do
  [ $n -le $G ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 1000 --warmup 10 > $O/bench_n$n.json 2> $O/bench_n$n.err
  python -c "import json,sys; j=json.loads(open('$O/bench_n$n.json').read().strip().splitlines()[-1]); print('N=%d value %.4g  e2e %.4g  gather_ms %.3f' % (j['n_gpus'], j['value'], j['e2e']['value'], j.get('gather_ms',0)))"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29600 profiles/config4_ensemble.py 100 > $O/config4_n$G.json 2> $O/config4.err; tail -1 $O/config4_n$G.json
