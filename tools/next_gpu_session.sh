#!/bin/bash
# What the round-2 GPU budget no longer covered, in the order it should be measured (one B200 unless noted):
#   gpurun --timeout 900 -- 'bash tools/next_gpu_session.sh'
# 1. wide wgrad stages / 5-d boxes BELOW their default thresholds (1 MB and 512 KB planes = the 1024^2-level layers at
#    N=2 / N=4, where the default falls back to 3-d boxes): SPC_WG_WIDE=1 SPC_WG_PAIR_WIDE=1 SPC_PW_BOX5=3 on tile shapes
#    of N=2/4/8 (tools/halo_cost_probe.py N prints fprop / wgrad per halo layer; tools/wgrad_probe.py takes square shapes)
# 2. bench at N=8 with the final binary (gpurun --gpus 8 -- 'bash tools/gpu_multi_bench_only.sh 8'); round 2's last N=8
#    record (27.9 ms) predates the 5-d boxes
# 3. ncu --set full of the wide pair wgrad and of pw_gemm with 5-d boxes (dram bytes, translation stalls) -> profiles/
# 4. next kernels: CTA-pair (cta_group::2) fprop/dgrad for the 1024^2 compute-heavy layers (they are L2->SM bound on
#    re-streamed weights, DESIGN.md "Address translation"); a 4-row tail box so that 52-channel operands can use 5-d boxes;
#    a small-C stem kernel; kind::tf32 for fp32 storage
mkdir -p gpurun_out
for n in 2 4 8; do
  echo "== tile shapes of N=$n, defaults"; timeout 120 python tools/halo_cost_probe.py $n | tail -1
  echo "== tile shapes of N=$n, wide / 5-d forced"; SPC_WG_WIDE=1 SPC_WG_PAIR_WIDE=1 SPC_PW_BOX5=3 timeout 120 python tools/halo_cost_probe.py $n | tail -1
done
timeout 200 python tools/wgrad_probe.py --wide > gpurun_out/wgrad_wide.log 2>&1; tail -30 gpurun_out/wgrad_wide.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tests_gpu.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_n1.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'])"
