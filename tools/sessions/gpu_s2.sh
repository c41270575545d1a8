#!/bin/bash
# Round-2 GPU call 2: first run of conv_tap.cu (guarded by timeouts: a hang must not take the box)
mkdir -p gpurun_out
timeout 120 python tools/tap_probe.py --quick > gpurun_out/r2_tap_quick.log 2>&1; echo "tap quick rc=$?"; tail -8 gpurun_out/r2_tap_quick.log
timeout 240 python tools/tap_probe.py > gpurun_out/r2_tap_probe.log 2>&1; echo "tap probe rc=$?"; tail -24 gpurun_out/r2_tap_probe.log
timeout 150 python tools/wgrad_probe.py --pair > gpurun_out/r2_wgrad_pair_all.log 2>&1; echo "pair all rc=$?"; grep -v "^$" gpurun_out/r2_wgrad_pair_all.log | tail -60
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q -x > gpurun_out/r2_tests_tap.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_tests_tap.log
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print(d['ms_per_step'],d['value'],d['e2e']['value'],d['config']['launch_mode'])
c=d.get('cudnn_baseline') or {}
print({k:(v['ms_per_step'] if isinstance(v,dict) and 'ms_per_step' in v else v) for k,v in c.items() if k!='what'})
print(d.get('model_stage'))
PY
