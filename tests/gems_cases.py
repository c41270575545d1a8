"""Shared driver for the GEMS-master trainer fixtures: the same worker runs either the UNMODIFIED
reference (tools/gen_trainer_golden.py, build container only) or this repo's mirrors
(tests/test_trainers.py) on CPU/gloo, selected by `impl`."""
import os
from types import SimpleNamespace

import torch
import torch.distributed as dist
import torch.nn as nn

IMG, IMG_SEQ = 16, 8

GEMS_CASES = [
    # layer-parallel GEMS master: 2 mirrored replicas of a 2-stage pipeline on 2 ranks
    dict(name="gems_lp2", kind="lp", world=2, split=2, balance=[4, 3], parts=1, batch=2, replications=1),
    dict(name="gems_lp2_parts2_rep2", kind="lp", world=2, split=2, balance=[4, 3], parts=2, batch=2, replications=2),
    # GEMS master + spatial parallelism: tiles of replica 1 on ranks 0,1 and of replica 2 on ranks 3,2
    dict(name="gems_sp2_vertical", kind="sp", world=4, P=2, split=3, balance=[4, 1, 2], parts=1, batch=2, replications=1,
         slice="vertical"),
    # same, with the parameter / gradient shipping protocol of --enable-master-comm-opt
    dict(name="gems_sp2_vertical_commopt", kind="sp", world=4, P=2, split=3, balance=[4, 1, 2], parts=1, batch=2,
         replications=1, slice="vertical", commopt=True),
]


def build_model(img, seed):
    torch.manual_seed(seed)
    return nn.Sequential(
        nn.Conv2d(3, 8, 1), nn.ReLU(), nn.Conv2d(8, 8, 1), nn.ReLU(),
        nn.Conv2d(8, 4, 3, stride=2, padding=1), nn.Flatten(), nn.Linear(4 * (img // 2) ** 2, 10))


def _mods(impl):
    if impl == "reference":
        from torchgems import gems_master, mp_pipeline, train_spatial, train_spatial_master
    else:
        from mpi4dl_b200.torchgems import gems_master, mp_pipeline, train_spatial, train_spatial_master
    return gems_master, mp_pipeline, train_spatial, train_spatial_master


def worker(rank, case, port, q, impl):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(case["world"]))
    dist.init_process_group("gloo", rank=rank, world_size=case["world"])
    torch.set_num_threads(1)
    gm, mp_, ts, tsm = _mods(impl)
    world, mb = case["world"], case["batch"] // case["parts"]
    nb = 2 * case["replications"]
    first, second = rank, world - 1 - rank                 # positions on the two rank lines
    if case["kind"] == "lp":
        gens = []
        for seed, pos in ((11, first), (22, second)):
            g = mp_.model_generator(model=build_model(IMG, seed), split_size=case["split"], input_size=(mb, 3, IMG, IMG),
                                    balance=case["balance"])
            g.ready_model(split_rank=pos, GET_SHAPES_ON_CUDA=False)
            gens.append(g)
        tm = gm.train_model_master(gens[0], gens[1], first, case["batch"], epochs=1, parts=case["parts"], ASYNC=True,
                                   replications=case["replications"])
        split = lambda x: x                                                           # noqa: E731
    else:
        P = case["P"]
        stage = lambda pos: pos // P if pos < P else pos - P + 1                      # noqa: E731
        seq = mp_.model_generator(model=build_model(IMG_SEQ, 1), split_size=case["split"], input_size=(mb, 3, IMG_SEQ, IMG_SEQ),
                                  balance=case["balance"])
        seq.get_output_shapes(False)
        shapes = ts.get_shapes_spatial(seq.shape_list, case["slice"], 1, [P], IMG // IMG_SEQ)
        gens = []
        for seed, pos in ((11, first), (22, second)):
            g = mp_.model_generator(model=build_model(IMG, seed), split_size=case["split"], input_size=(mb, 3, IMG, IMG),
                                    balance=case["balance"], shape_list=shapes)
            g.ready_model(split_rank=stage(pos))
            gens.append(g)
        c1 = SimpleNamespace(mp_size=world, local_rank=first)
        c2 = SimpleNamespace(mp_size=world, local_rank=second)
        tm = tsm.train_spatial_model_master(gens[0], gens[1], case["batch"], 1, P, case["slice"], c1, c2, LOCAL_DP_LP=1,
                                            parts=case["parts"], ASYNC=True, replications=case["replications"])

        def split(x):
            if first < P:
                return ts.split_input(x, IMG, case["slice"], first, [P])
            if second < P:
                return ts.split_input(x, IMG, case["slice"], second, [P])
            return x
    losses = []
    for step in range(3):
        g = torch.Generator().manual_seed(300 + step)
        x = torch.randn(nb * case["batch"], 3, IMG, IMG, generator=g)
        y = torch.randint(0, 10, (nb * case["batch"],), generator=g)
        if case.get("commopt"):
            loss, _ = tm.run_step_allreduce(split(x), y, step % 2 == 1)
            (tm.train_model1 if step % 2 == 1 else tm.train_model2).update()
        else:
            loss, _ = tm.run_step(split(x), y)
            tm.train_model1.update()
            tm.train_model2.update()
        losses.append(float(loss))
    q.put((rank, losses))
    dist.barrier()
    dist.destroy_process_group()
