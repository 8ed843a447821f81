#!/bin/bash
# round 2, pass k (N GPUs of one box; N = $1): C-ABI ensembles (tests + the C++ host on BASELINE configs[3]) and the bench line under torchrun
N=${1:-2}
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/gpus_n$N.txt
( time timeout 900 python -m pytest tests -m gpu -x -q -k "ensemble" ) > $O/pytest_ensemble_n$N.log 2>&1; tail -3 $O/pytest_ensemble_n$N.log
g++ -std=c++17 -O1 tests/ensemble_main.cpp -o /tmp/ensemble_main -I/usr/local/cuda/include -Lhamilton_b200/lib -lhamilton_b200 -Wl,-rpath,$PWD/hamilton_b200/lib -L/usr/local/cuda/lib64 -lcudart
for nd in 1 $N; do
  timeout 600 /tmp/ensemble_main $nd 6 8388608 1000 | tee -a $O/ensemble_cpp_config4.txt
done
timeout 600 /tmp/ensemble_main $N 6 1048576 50 check | tee -a $O/ensemble_cpp_config4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; tail -c 400 $O/bench_n$N.json; tail -2 $O/bench_n$N.err
