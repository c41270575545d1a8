#!/bin/bash
# short multi-GPU check after the boundary-GEMM fix-up: bench at N (graph replay), and at N=8 config 4 again
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node $N --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err
echo "bench rc=$?"; grep "bench " gpurun_out/r2i_bench_n$N.err | tail -3
python -c "
import json;d=json.load(open('gpurun_out/r2i_bench_n$N.json'));print('N=$N', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['launch_mode'], d['gpu_launches'])"
if [ "$N" = "8" ]; then
  timeout 300 $TR --nproc-per-node 8 --master-port 29712 benchmarks/gems_master_with_spatial_parallelism/benchmark_amoebanet_gems_master_with_sp.py \
     --image-size 2048 --num-spatial-parts 4 --slice-method square --split-size 5 --spatial-size 1 --batch-size 1 --times 2 \
     --num-layers 18 --num-filters 416 --dtype bf16 --steps 6 > gpurun_out/r2_cfg4_gems_sp.log 2>&1; echo "cfg4 rc=$?"; grep -E "images per sec|Mean|LOSS|Error|error" gpurun_out/r2_cfg4_gems_sp.log | tail -8
fi
