"""Loss sequences of the UNMODIFIED reference's layer-parallel trainer (mp_pipeline.train_model) on
CPU/gloo -> tests/golden/trainer_golden.json.  Run in the build container only."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "trainer_golden.json")


def build_model():
    torch.manual_seed(1234)
    return nn.Sequential(
        nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, stride=2, padding=1), nn.ReLU(),
        nn.Conv2d(8, 4, 3, padding=1), nn.Flatten(), nn.Linear(4 * 8 * 8, 10))


def data(step, batch):
    g = torch.Generator().manual_seed(100 + step)
    return torch.randn(batch, 3, 16, 16, generator=g), torch.randint(0, 10, (batch,), generator=g)


CASES = [dict(name="lp2_parts1", world=2, split=2, parts=1, batch=4, balance=None),
         dict(name="lp3_parts2", world=3, split=3, parts=2, batch=4, balance=[2, 3, 2])]


def build_sp_model(img):
    """Spatial stage of tile-local layers (1x1 convs need no halo, so plain nn.Conv2d is exact on a
    tile), then an ordinary tail that needs the stitched map."""
    torch.manual_seed(4321)
    return nn.Sequential(
        nn.Conv2d(3, 8, 1), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.ReLU(),
        nn.Conv2d(8, 4, 3, stride=2, padding=1), nn.Flatten(), nn.Linear(4 * (img // 2) ** 2, 10))


SP_CASES = [
    dict(name="sp2_vertical", P=2, spatial_size=1, split=2, balance=[4, 3], parts=1, batch=2, slice="vertical", inverse=False),
    dict(name="sp4_square_parts2", P=4, spatial_size=1, split=3, balance=[4, 1, 2], parts=2, batch=4, slice="square", inverse=False),
    dict(name="sp2x2_horizontal", P=2, spatial_size=2, split=3, balance=[2, 2, 3], parts=1, batch=2, slice="horizontal", inverse=False),
    dict(name="sp2_vertical_inverse", P=2, spatial_size=1, split=2, balance=[4, 3], parts=1, batch=2, slice="vertical", inverse=True),
]
IMG, IMG_SEQ = 16, 8


def sp_world(case):
    return case["P"] * case["spatial_size"] + case["split"] - case["spatial_size"]


def sp_worker(rank, case, port, q):
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.install()
    world = sp_world(case)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from types import SimpleNamespace
    from torchgems.mp_pipeline import model_generator
    from torchgems.train_spatial import get_shapes_spatial, split_input, train_model_spatial
    P, S = case["P"], case["spatial_size"]
    nsp_list = [P] * S
    nsp = P if S == 1 else nsp_list
    local_rank = world - 1 - rank if case["inverse"] else rank          # position on the rank line
    split_rank = local_rank // P if local_rank < P * S else local_rank - P * S + S
    mb = case["batch"] // case["parts"]
    seq = model_generator(model=build_sp_model(IMG_SEQ), split_size=case["split"], input_size=(mb, 3, IMG_SEQ, IMG_SEQ),
                          balance=case["balance"])
    seq.ready_model(split_rank=split_rank, GET_SHAPES_ON_CUDA=False)
    shapes = get_shapes_spatial(seq.shape_list, case["slice"], S, nsp_list, IMG // IMG_SEQ)
    gen = model_generator(model=build_sp_model(IMG), split_size=case["split"], input_size=(mb, 3, IMG, IMG),
                          balance=case["balance"], shape_list=shapes)
    gen.ready_model(split_rank=split_rank)
    tm = train_model_spatial(gen, local_rank, case["batch"], epochs=1, spatial_size=S, num_spatial_parts=nsp, parts=case["parts"],
                             ASYNC=True, GEMS_INVERSE=case["inverse"], slice_method=case["slice"],
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


def gems_worker(rank, case, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(HERE, "..", "tests"))
    import ref_shim
    ref_shim.install()
    import gems_cases
    gems_cases.worker(rank, case, port, q, "reference")


def worker(rank, case, port, q):
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.install()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(case["world"]))
    dist.init_process_group("gloo", rank=rank, world_size=case["world"])
    torch.set_num_threads(1)
    from torchgems.mp_pipeline import model_generator, train_model
    model = build_model()
    mb = case["batch"] // case["parts"]
    gen = model_generator(model=model, split_size=case["split"], input_size=(mb, 3, 16, 16), balance=case["balance"])
    gen.ready_model(split_rank=rank, GET_SHAPES_ON_CUDA=False)
    tm = train_model(gen, rank, batch_size=case["batch"], epochs=1, criterion=None, optimizer=None, parts=case["parts"], ASYNC=True)
    losses = []
    for step in range(3):
        x, y = data(step, case["batch"])
        loss, _ = tm.run_step(x, y)
        tm.update()
        losses.append(float(loss))
    q.put((rank, losses, [list(s) if not isinstance(s, list) else [list(t) for t in s] for s in gen.shape_list]))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ctx = mp.get_context("spawn")
    res = {}
    port = 29870
    for case in CASES:
        port += 1
        q = ctx.SimpleQueue()
        ps = [ctx.Process(target=worker, args=(r, case, port, q)) for r in range(case["world"])]
        for p in ps:
            p.start()
        got = {r: (l, s) for r, l, s in (q.get() for _ in ps)}
        for p in ps:
            p.join()
        res[case["name"]] = dict(case=case, losses=got[case["world"] - 1][0], shape_list=got[0][1])
        print(case["name"], got[case["world"] - 1][0], flush=True)
    sp = {}
    for case in SP_CASES:
        port += 1
        world = sp_world(case)
        q = ctx.SimpleQueue()
        ps = [ctx.Process(target=sp_worker, args=(r, case, port, q)) for r in range(world)]
        for p in ps:
            p.start()
        got = {r: (l, s) for r, l, s in (q.get() for _ in ps)}
        for p in ps:
            p.join()
        sp[case["name"]] = dict(case=case, losses=got[world - 1][0], shape_list=got[0][1])
        print(case["name"], got[world - 1][0], flush=True)
    res_sp = sp
    sys.path.insert(0, os.path.join(HERE, "..", "tests"))
    import gems_cases
    res_gems = {}
    for case in gems_cases.GEMS_CASES:
        port += 1
        q = ctx.SimpleQueue()
        ps = [ctx.Process(target=gems_worker, args=(r, case, port, q)) for r in range(case["world"])]
        for p in ps:
            p.start()
        got = dict(q.get() for _ in ps)
        for p in ps:
            p.join()
        res_gems[case["name"]] = dict(case=case, losses={str(r): l for r, l in sorted(got.items())})
        print(case["name"], got, flush=True)
    json.dump({"source": "tools/gen_trainer_golden.py on unmodified /root/reference mp_pipeline.py (gloo, CPU)", "cases": res,
               "sp_source": "same script, unmodified /root/reference train_spatial.py (train_model_spatial, get_shapes_spatial, split_input)",
               "sp_cases": res_sp,
               "gems_source": "tests/gems_cases.py worker on unmodified /root/reference gems_master.py / train_spatial_master.py",
               "gems_cases": res_gems},
              open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
