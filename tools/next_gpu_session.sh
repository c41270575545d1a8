#!/bin/bash
# First GPU call of the next session (one B200, ~4 minutes): what round 1 could not measure any more.
#   gpurun --timeout 600 -- 'bash tools/next_gpu_session.sh'
# 1. the cta_group::2 wgrad pair kernel, never run on hardware yet (outer timeout: a hang must not take the box)
# 2. the single-CTA wgrad tiling sweep
# 3. the full GPU test suite + the default bench line
mkdir -p gpurun_out
timeout 90 python tools/wgrad_probe.py --pair --quick > gpurun_out/wgrad_pair.log 2>&1; echo "pair probe rc=$?"
tail -12 gpurun_out/wgrad_pair.log
timeout 240 python tools/wgrad_probe.py > gpurun_out/wgrad_sweep.log 2>&1; echo "sweep rc=$?"
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_gpu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
# 4. the self-checking halo benchmarks (reference's own validation tools), 4 tiles sharing the GPU over gloo
export SPCONV_DIST_BACKEND=gloo
for m in vertical square; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 \
    benchmarks/communication/halo/benchmark_sp_halo_exchange_conv.py --image-size 64 --halo-len 3 --num-spatial-parts 4 \
    --slice-method $m --in-channels 2 --out-channels 8 --iterations 10 --enable-val-recv-tensors --enable-val-conv 2>&1 | grep "Rank:"
done
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 \
  benchmarks/communication/halo/benchmark_sp_halo_exchange.py --image-size 32 --halo-len 2 --num-spatial-parts 2 --slice-method horizontal 2>&1 | grep "Rank:"
