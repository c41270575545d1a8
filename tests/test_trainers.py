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
_G = json.load(open(os.path.join(ROOT, "tests", "golden", "trainer_golden.json")))
GOLD = _G["cases"]
SP_GOLD = _G["sp_cases"]
GEMS_GOLD = _G["gems_cases"]
IMG, IMG_SEQ = 16, 8


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


def build_sp_model(img):
    torch.manual_seed(4321)
    return nn.Sequential(
        nn.Conv2d(3, 8, 1), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.ReLU(),
        nn.Conv2d(8, 4, 3, stride=2, padding=1), nn.Flatten(), nn.Linear(4 * (img // 2) ** 2, 10))


def _sp_world(case):
    return case["P"] * case["spatial_size"] + case["split"] - case["spatial_size"]


def _sp_worker(rank, case, port, q):
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, ROOT)
    world = _sp_world(case)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      CUDA_VISIBLE_DEVICES="")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from mpi4dl_b200.torchgems.mp_pipeline import model_generator
    from mpi4dl_b200.torchgems.train_spatial import get_shapes_spatial, split_input, train_model_spatial
    P, S = case["P"], case["spatial_size"]
    nsp_list = [P] * S
    nsp = P if S == 1 else nsp_list
    local_rank = world - 1 - rank if case["inverse"] else rank
    split_rank = local_rank // P if local_rank < P * S else local_rank - P * S + S
    mb = case["batch"] // case["parts"]
    seq = model_generator(model=build_sp_model(IMG_SEQ), split_size=case["split"], input_size=(mb, 3, IMG_SEQ, IMG_SEQ),
                          balance=case["balance"])
    seq.ready_model(split_rank=split_rank, GET_SHAPES_ON_CUDA=False)
    shapes = get_shapes_spatial(seq.shape_list, case["slice"], S, nsp_list, IMG // IMG_SEQ)
    gen = model_generator(model=build_sp_model(IMG), split_size=case["split"], input_size=(mb, 3, IMG, IMG),
                          balance=case["balance"], shape_list=shapes)
    gen.ready_model(split_rank=split_rank)
    tm = train_model_spatial(gen, local_rank, case["batch"], epochs=1, spatial_size=S, num_spatial_parts=nsp,
                             parts=case["parts"], ASYNC=True, GEMS_INVERSE=case["inverse"], slice_method=case["slice"],
                             mpi_comm=SimpleNamespace(mp_size=world))
    losses = []
    for step in range(3):
        g = torch.Generator().manual_seed(200 + step)
        x = torch.randn(case["batch"], 3, IMG, IMG, generator=g)
        y = torch.randint(0, 10, (case["batch"],), generator=g)
        if local_rank < P:
            x = split_input(x, IMG, case["slice"], local_rank, nsp_list)
        loss, _ = tm.run_step(x, y)
        tm.update()
        losses.append(float(loss))
    q.put((local_rank, losses, [list(s) if not isinstance(s, list) else [list(t) for t in s] for s in shapes]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("idx,name", list(enumerate(sorted(SP_GOLD))))
def test_sp_trainer_matches_reference_losses(idx, name):
    """SP+LP trainer (tiles -> join rank -> tail), incl. two spatial stages, square/strip slicing,
    micro-batches and the mirrored (GEMS inverse) rank line, against the reference's own losses."""
    case = SP_GOLD[name]["case"]
    world = _sp_world(case)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    ps = [ctx.Process(target=_sp_worker, args=(r, case, 29890 + idx, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = {r: (l, s) for r, l, s in (q.get() for _ in ps)}
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == SP_GOLD[name]["shape_list"]
    assert got[world - 1][0] == pytest.approx(SP_GOLD[name]["losses"], rel=1e-6, abs=1e-6)


def _gems_worker(rank, case, port, q, zero_init):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    os.environ["SPCONV_GEMS_REFERENCE_ZERO_INIT"] = "1" if zero_init else "0"
    os.environ["SPCONV_REFERENCE_ZERO_GRAD"] = "1" if zero_init else "0"   # the reference's zero_grad() (drops .grad)
    import gems_cases
    gems_cases.worker(rank, case, port, q, "ours")


def _run_gems(case, port, zero_init):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    ps = [ctx.Process(target=_gems_worker, args=(r, case, port, q, zero_init)) for r in range(case["world"])]
    for p in ps:
        p.start()
    got = dict(q.get() for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    return got


@pytest.mark.parametrize("idx,name", list(enumerate(sorted(GEMS_GOLD))))
def test_gems_master_trainers_match_reference_losses(idx, name):
    """GEMS master (two mirrored replicas) on LP and on SP+LP: per-rank loss sequences of the
    reference.  The SP variant runs with the reference's zero-initialised flat parameter buffers
    (see train_spatial_master._flatten) to compare like with like."""
    case = GEMS_GOLD[name]["case"]
    got = _run_gems(case, 29900 + idx, zero_init=True)
    for r, want in GEMS_GOLD[name]["losses"].items():
        assert got[int(r)] == pytest.approx(want, rel=1e-6, abs=1e-6), r


def test_gems_sp_master_keeps_initial_parameters_by_default():
    """Default (non-reference) behaviour: flattening keeps the initial weights, so the two replicas
    start from their own initialisation and the first loss is not ln(10)."""
    case = GEMS_GOLD["gems_sp2_vertical"]["case"]
    got = _run_gems(case, 29910, zero_init=False)
    import math
    first = got[3][0]
    assert abs(first - math.log(10)) > 1e-3 and math.isfinite(first)


def test_update_keeps_flat_gradient_aliases():
    """after train_model.update() the parameters' .grad must still be views of the flat gradient
    buffer train_spatial_model_master ships between mirror ranks (zero_grad must not drop them)."""
    import torch.nn as nn
    from mpi4dl_b200.torchgems import mp_pipeline
    from mpi4dl_b200.torchgems.train_spatial_master import train_spatial_model_master as M

    os.environ.pop("SPCONV_REFERENCE_ZERO_GRAD", None)
    model = nn.Sequential(nn.Linear(4, 3), nn.Linear(3, 2))
    holder = M.__new__(M)
    holder.device = torch.device("cpu")
    size = sum(p.numel() for p in model.parameters())
    flat_p, flat_g = M._flatten(holder, model, size)
    tm = mp_pipeline.train_model.__new__(mp_pipeline.train_model)
    tm.optimizer = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(2):
        model(torch.randn(5, 4)).sum().backward()
        assert flat_g.abs().sum() > 0                       # backward accumulated INTO the flat buffer
        tm.update()
        assert float(flat_g.abs().sum()) == 0.0             # zeroed in place
        off = 0
        for p_ in model.parameters():
            assert p_.grad is not None and p_.grad.data_ptr() == flat_g[off:].data_ptr(), step
            off += p_.numel()


def test_spatial_master_config():
    from mpi4dl_b200.torchgems.train_spatial_master import verify_spatial_master_config
    verify_spatial_master_config("square", 1024, [4], 1, 8)
    with pytest.raises(AssertionError):
        verify_spatial_master_config("square", 1024, [4], 1, 5)     # replica tiles would share GPUs


def test_spatial_config_helpers():
    from mpi4dl_b200.torchgems.train_spatial import get_shapes_spatial, split_input, verify_spatial_config
    verify_spatial_config("square", 8192, [4])
    verify_spatial_config("vertical", 1024, [8, 8])
    for bad in (("diagonal", 1024, [4]), ("square", 1000, [4]), ("vertical", 1024, [3]), ("vertical", 1024, [4, 2])):
        with pytest.raises(AssertionError):
            verify_spatial_config(*bad)
    shapes = [(2, 8, 32, 32), [(2, 8, 16, 16), (2, 4, 16, 16)], (2, 16, 8, 8), (2, 10)]
    assert get_shapes_spatial(shapes, "square", 2, [4, 4], 4) == \
        [(2, 8, 64, 64), [(2, 8, 32, 32), (2, 4, 32, 32)], (2, 16, 32, 32), (2, 10)]
    assert get_shapes_spatial(shapes, "vertical", 1, [4], 2) == \
        [(2, 8, 64, 16), [(2, 8, 32, 32), (2, 4, 32, 32)], (2, 16, 16, 16), (2, 10)]
    assert get_shapes_spatial(shapes, "horizontal", 2, [2, 2], 1)[:2] == [(2, 8, 16, 32), [(2, 8, 8, 16), (2, 4, 8, 16)]]
    x = torch.arange(2 * 1 * 8 * 8).reshape(2, 1, 8, 8)
    assert torch.equal(split_input(x, 8, "square", 3, [4]), x[:, :, 4:, 4:])
    assert torch.equal(split_input(x, 8, "square", 1, [4]), x[:, :, :4, 4:])
    assert torch.equal(split_input(x, 8, "vertical", 1, [4]), x[:, :, :, 2:4])
    assert torch.equal(split_input(x, 8, "horizontal", 3, [4]), x[:, :, 6:, :])
    # odd tile counts: the last strip takes the remainder (train_spatial.py:268-276)
    assert split_input(x, 8, "vertical", 2, [3]).shape[-1] == 4


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
