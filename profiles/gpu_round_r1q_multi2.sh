#!/bin/bash
# 2 GPUs: NUMA-bound ranks in the e2e leg, gather without count exchange; host topology
mkdir -p gpurun_out/r1q
O=gpurun_out/r1q
nvidia-smi topo -m > $O/topo.txt 2>&1; head -14 $O/topo.txt; nproc; numactl -H 2>/dev/null | head -5; lscpu | grep -i -E "numa|socket|model name" | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 300 --warmup 10 > $O/bench_n2.json 2> $O/bench_n2.err
python -c "import json,sys; j=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1]); print('N=%d value %.4g  e2e %.4g  gather_ms %.3f' % (j['n_gpus'], j['value'], j['e2e']['value'], j.get('gather_ms',0)), j['e2e'].get('host_binding'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29600 profiles/config4_ensemble.py 100 > $O/config4_n2.json 2> $O/config4.err; tail -1 $O/config4_n2.json
