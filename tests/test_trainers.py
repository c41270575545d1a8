"""CPU-only (gloo): the layer-parallel trainer mirror (mpi4dl_b200.torchgems.mp_pipeline) reproduces the
loss sequences of the UNMODIFIED reference trainer (tools/gen_trainer_golden.py ->
tests/golden/trainer_golden.json); parser / utils mirrors keep the reference's flag set."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "trainer_golden.json")))["cases"]


def build_model():
    torch.manual_seed(1234)
    return nn.Sequential(
        nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, stride=2, padding=1), nn.ReLU(),
        nn.Conv2d(8, 4, 3, padding=1), nn.Flatten(), nn.Linear(4 * 8 * 8, 10))


def data(step, batch):
    g = torch.Generator().manual_seed(100 + step)
    return torch.randn(batch, 3, 16, 16, generator=g), torch.randint(0, 10, (batch,), generator=g)


def _worker(rank, case, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(case["world"]),
                      CUDA_VISIBLE_DEVICES="")
    dist.init_process_group("gloo", rank=rank, world_size=case["world"])
    torch.set_num_threads(1)
    from mpi4dl_b200.torchgems.mp_pipeline import model_generator, train_model
    model = build_model()
    mb = case["batch"] // case["parts"]
    gen = model_generator(model=model, split_size=case["split"], input_size=(mb, 3, 16, 16), balance=case["balance"])
    gen.ready_model(split_rank=rank, GET_SHAPES_ON_CUDA=False)
    tm = train_model(gen, rank, batch_size=case["batch"], epochs=1, parts=case["parts"], ASYNC=True)
    losses = []
    for step in range(3):
        x, y = data(step, case["batch"])
        loss, _ = tm.run_step(x, y)
        tm.update()
        losses.append(float(loss))
    shapes = [list(s) if not isinstance(s, list) else [list(t) for t in s] for s in gen.shape_list]
    q.put((rank, losses, shapes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("idx,name", list(enumerate(sorted(GOLD))))
def test_lp_trainer_matches_reference_losses(idx, name):
    case = GOLD[name]["case"]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    ps = [ctx.Process(target=_worker, args=(r, case, 29880 + idx, q)) for r in range(case["world"])]
    for p in ps:
        p.start()
    got = {r: (l, s) for r, l, s in (q.get() for _ in ps)}
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == GOLD[name]["shape_list"]
    assert got[case["world"] - 1][0] == pytest.approx(GOLD[name]["losses"], rel=1e-6, abs=1e-6)


def test_parser_namespace_matches_reference():
    from mpi4dl_b200.torchgems import parser
    ns = vars(parser.get_parser().parse_args([]))
    # defaults of the reference's flag set (src/torchgems/parser.py:21-143)
    assert ns == {"verbose": False, "batch_size": 32, "parts": 1, "split_size": 2, "num_spatial_parts": "4", "spatial_size": 1,
                  "times": 1, "image_size": 32, "num_epochs": 1, "num_layers": 18, "num_filters": 416, "num_classes": 10,
                  "balance": None, "halo_d2": False, "fused_layers": 1, "local_DP": 1, "slice_method": "square", "app": 3,
                  "datapath": "./train", "enable_master_comm_opt": False, "num_workers": 0}
    a = parser.get_parser().parse_args("--image-size 8192 --num-spatial-parts 4 --split-size 4 --halo-D2 --local-DP 2".split())
    assert (a.image_size, a.num_spatial_parts, a.split_size, a.halo_d2, a.local_DP) == (8192, "4", 4, True, 2)


def test_utils():
    from mpi4dl_b200.torchgems.utils import get_depth, isPowerTwo
    assert isPowerTwo(8192) and isPowerTwo(1) and not isPowerTwo(24)
    assert get_depth(2, 11) == 101 and get_depth(1, 3) == 20
