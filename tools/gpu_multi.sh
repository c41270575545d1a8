#!/bin/bash
# Multi-GPU session (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): what BASELINE.json configs 3-5 ask to RUN.
#   N=2,4: halo sweep (config 5), bench.py at N, transport tests on separate GPUs
#   N=8  : halo sweep P=8, bench at N=8, SP+LP AmoebaNet (config 3, 7 ranks), GEMS-MASTER+SP (config 4, 8 ranks),
#          ResNet-101 SP (config 2: 4 tiles + 1 LP rank = 5 ranks)
N=${1:-2}
mkdir -p gpurun_out profiles
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== halo sweep P=$N"
SW="--tiles 2048 4096 8192 16384"; [ "$N" = "8" ] && SW="--tiles 2048 8192 --slice-methods vertical"
timeout 300 $TR --nproc-per-node $N --master-port 29701 benchmarks/communication/halo/halo_sweep.py --reference-point $SW --iterations 30 \
   --out gpurun_out/r2_halo_sweep_P$N.json > gpurun_out/r2_halo_sweep_P$N.log 2>&1; echo "sweep rc=$?"; grep '^{' gpurun_out/r2_halo_sweep_P$N.log | head -40
echo "== reference self-checking halo scripts on $N GPUs"
M=vertical; [ "$N" = "4" ] && M=square
# (known answers are exact integers only while the 7x7 box sums stay below 2^24: validate at 128^2, time at 1024^2)
timeout 120 $TR --nproc-per-node $N --master-port 29702 benchmarks/communication/halo/benchmark_sp_halo_exchange_conv.py --image-size 128 \
   --halo-len 3 --num-spatial-parts $N --slice-method $M --in-channels 1 --out-channels 256 --iterations 20 \
   --enable-val-recv-tensors --enable-val-conv 2>&1 | grep "Rank:" | sort | head -20
timeout 120 $TR --nproc-per-node $N --master-port 29706 benchmarks/communication/halo/benchmark_sp_halo_exchange_with_compute.py --image-size 1024 \
   --halo-len 3 --num-spatial-parts $N --slice-method $M --iterations 100 2>&1 | grep "Rank:" | sort | head -20
timeout 120 $TR --nproc-per-node $N --master-port 29703 benchmarks/communication/halo/benchmark_sp_halo_exchange.py --image-size 1024 \
   --halo-len 3 --num-spatial-parts $N --slice-method vertical 2>&1 | grep "Rank:" | sort | head -20
echo "== bench N=$N"
timeout 300 $TR --nproc-per-node $N --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?"; grep "bench " gpurun_out/r2_bench_n$N.err | tail -6
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n$N.json'));print('N=$N', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['launch_mode'], d['gpu_launches'])"
timeout 200 $TR --nproc-per-node $N --master-port 29705 bench.py --gpus $N --steps 10 --warmup 3 --graph off > gpurun_out/r2_bench_n${N}_eager.json 2> gpurun_out/r2_bench_n${N}_eager.err
python -c "
import json;d=json.load(open('gpurun_out/r2_bench_n${N}_eager.json'));print('N=$N eager', d['ms_per_step'], d['value'], d['e2e']['value'])"
if [ "$N" = "2" ]; then
  timeout 300 python -m pytest tests/test_gpu_peer_transport.py tests/test_gpu_gems_sp.py -m gpu -q > gpurun_out/r2_tests_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_tests_n2.log
fi
if [ "$N" = "8" ]; then
  echo "== config 3: AmoebaNet-D SP+LP, split_size 4, 4 tiles + 3 LP ranks"
  timeout 400 $TR --nproc-per-node 7 --master-port 29711 benchmarks/spatial_parallelism/benchmark_amoebanet_sp.py --image-size 4096 \
     --num-spatial-parts 4 --slice-method square --split-size 4 --spatial-size 1 --batch-size 1 --num-layers 18 --num-filters 416 \
     --dtype bf16 --steps 6 > gpurun_out/r2_cfg3_amoebanet_sp_lp.log 2>&1; echo "cfg3 rc=$?"; grep -E "images per sec|Mean|LOSS|Error|error" gpurun_out/r2_cfg3_amoebanet_sp_lp.log | tail -8
  echo "== config 4: AmoebaNet-D GEMS-MASTER+SP, split_size 5, 8 ranks"
  timeout 400 $TR --nproc-per-node 8 --master-port 29712 benchmarks/gems_master_with_spatial_parallelism/benchmark_amoebanet_gems_master_with_sp.py \
     --image-size 2048 --num-spatial-parts 4 --slice-method square --split-size 5 --spatial-size 1 --batch-size 1 --times 2 \
     --num-layers 18 --num-filters 416 --dtype bf16 --steps 6 > gpurun_out/r2_cfg4_gems_sp.log 2>&1; echo "cfg4 rc=$?"; grep -E "images per sec|Mean|LOSS|Error|error" gpurun_out/r2_cfg4_gems_sp.log | tail -8
  echo "== config 2: ResNet-v2 SP (4 tiles + 1 LP rank), 4096^2 bf16"
  timeout 400 $TR --nproc-per-node 5 --master-port 29713 benchmarks/spatial_parallelism/benchmark_resnet_sp.py --image-size 4096 \
     --num-spatial-parts 4 --slice-method square --split-size 2 --spatial-size 1 --batch-size 1 --dtype bf16 --steps 6 \
     > gpurun_out/r2_cfg2_resnet_sp.log 2>&1; echo "cfg2 rc=$?"; grep -E "images per sec|Mean|LOSS|Error|error" gpurun_out/r2_cfg2_resnet_sp.log | tail -8
fi
