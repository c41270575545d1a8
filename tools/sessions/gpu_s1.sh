#!/bin/bash
# Round-2 GPU call 1 (one B200): full GPU suite incl. the new full-size parity test, the cta_group::2 wgrad pair
# kernel's first run, and the default bench line with the cuDNN comparator + CUDA-graph replay.
mkdir -p gpurun_out
timeout 90 python tools/wgrad_probe.py --pair --quick > gpurun_out/r2_wgrad_pair.log 2>&1; echo "pair probe rc=$?"
tail -15 gpurun_out/r2_wgrad_pair.log
timeout 1100 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r2_tests_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2_tests_gpu.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['launch_mode'])
c=d.get('cudnn_baseline') or {}
print({k:(v['ms_per_step'] if isinstance(v,dict) and 'ms_per_step' in v else v) for k,v in c.items() if k!='what'})
print(d.get('model_stage'))
PY
