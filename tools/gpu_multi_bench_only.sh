#!/bin/bash
# bench at N GPUs only (graph replay)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node $N --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err
echo "bench rc=$?"; grep "bench " gpurun_out/r2j_bench_n$N.err | tail -2
