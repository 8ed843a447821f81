#!/bin/bash
# round 2, pass m (8 GPUs): is the host the ceiling of the multi-GPU e2e leg?  copy engines only, all ranks at once vs one rank;
# the e2e leg with staged copies instead of the zero-copy kernels
N=${1:-8}
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 profiles/exp/exp_pcie_multi.py 2>/dev/null | grep -v "^NCCL" | tee $O/pcie_multi_n$N.txt
HB_HOST_DIRECT=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 3 --no-extras 2>/dev/null | grep "^{" > $O/bench_n${N}_staged.json
python -c "
import json;d=json.load(open('$O/bench_n${N}_staged.json'));print('e2e with staged copies (HB_HOST_DIRECT=0), $N ranks: %.4g steps/s' % d['e2e']['value'])" | tee -a $O/pcie_multi_n$N.txt
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" | tee -a $O/pcie_multi_n$N.txt
nvidia-smi topo -m 2>/dev/null | head -14 >> $O/pcie_multi_n$N.txt
