"""CPU smoke of the layer-parallel and GEMS-master benchmark scripts (torchrun, gloo): the same
command lines as the reference's scripts, tiny shapes.  Includes a 4-process run with two
data-parallel replicas of a 2-stage pipeline -- the configuration whose peer addressing the
reference gets wrong (positions on the rank line used as process ranks)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

RUNS = [
    ("lp_resnet_2", 2, "layer_parallelism/benchmark_resnet_lp.py",
     "--split-size 2 --image-size 32 --batch-size 4 --parts 2 --steps 2"),
    ("lp_amoebanet_2x2_replicas", 4, "layer_parallelism/benchmark_amoebanet_lp.py",
     "--split-size 2 --image-size 64 --batch-size 2 --num-layers 3 --num-filters 64 --steps 2"),
    ("gems_resnet_2", 2, "gems_master_model/benchmark_resnet_gems_master.py",
     "--split-size 2 --image-size 32 --batch-size 2 --times 2 --steps 2"),
]


# Scripts with spatial layers: control flow only.  tests/cpu_smoke_hooks/sitecustomize.py (test infrastructure,
# activated by SPCONV_TEST_CPU_SMOKE=1 + PYTHONPATH) swaps the libspconv-backed layers for halo-less PyTorch ops,
# so rank arithmetic, trainers, collectives and script plumbing run on CPU/gloo; numerics are NOT checked here
# (tests/test_gpu_sp_trainer.py does that on the GPU with the real kernels).
SPATIAL_RUNS = [
    ("sp_amoebanet_d2_4tiles", 5, "spatial_parallelism/benchmark_amoebanet_sp.py",
     "--image-size 64 --num-spatial-parts 4 --slice-method square --split-size 2 --batch-size 1 --num-layers 6 "
     "--num-filters 64 --steps 2 --halo-D2"),
    ("sp_resnet_d2_2tiles", 3, "spatial_parallelism/benchmark_resnet_sp.py",
     "--image-size 64 --num-spatial-parts 2 --slice-method vertical --split-size 2 --batch-size 1 --steps 2 --halo-D2 "
     "--fused-layers 2"),
    ("gems_sp_resnet", 4, "gems_master_with_spatial_parallelism/benchmark_resnet_gems_master_with_sp.py",
     "--split-size 3 --num-spatial-parts 2 --slice-method vertical --image-size 64 --batch-size 2 --times 2 --steps 2"),
    ("gems_sp_resnet_commopt", 4, "gems_master_with_spatial_parallelism/benchmark_resnet_gems_master_with_sp.py",
     "--split-size 3 --num-spatial-parts 2 --slice-method vertical --image-size 64 --batch-size 2 --times 2 --steps 2 "
     "--enable-master-comm-opt"),
]


def _run(idx, nproc, script, flags, extra_env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(29750 + idx), os.path.join(ROOT, "benchmarks", script)] + flags.split()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=280,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1", **extra_env))
    assert out.returncode == 0, out.stderr[-3000:]
    assert "Mean " in out.stdout and "Global loss" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("idx,name,nproc,script,flags", [(i,) + r for i, r in enumerate(RUNS)], ids=[r[0] for r in RUNS])
def test_benchmark_script_runs(idx, name, nproc, script, flags):
    _run(idx, nproc, script, flags, {})


@pytest.mark.parametrize("idx,name,nproc,script,flags", [(10 + i,) + r for i, r in enumerate(SPATIAL_RUNS)],
                         ids=[r[0] for r in SPATIAL_RUNS])
def test_spatial_benchmark_script_control_flow(idx, name, nproc, script, flags):
    hooks = os.path.join(ROOT, "tests", "cpu_smoke_hooks")
    _run(idx, nproc, script, flags, {"SPCONV_TEST_CPU_SMOKE": "1",
                                     "PYTHONPATH": os.pathsep.join([hooks, ROOT, os.environ.get("PYTHONPATH", "")])})
