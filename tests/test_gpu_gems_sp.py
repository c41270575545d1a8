"""-m gpu: GEMS-MASTER + spatial parallelism end to end on the GPU (BASELINE config 4 in miniature):
two mirrored replicas of an SP+LP pipeline -- train_spatial_model_master over two train_model_spatial instances, the
second on the mirrored rank line (GEMS_INVERSE) -- whose spatial stages run conv_spatial / Pool on libspconv with
the halo mailboxes between the tile ranks of EACH replica.  World = 4 processes (replica 1: tiles on ranks 0,1,
join 2, tail 3; replica 2: tiles on ranks 3,2, join 1, tail 0); with fewer than 4 GPUs they share cuda:0 and
torch.distributed runs on gloo.  The step is the benchmark script's (benchmark_gems_master_with_sp.py): run_step, then
SyncAllreduce.apply_allreduce_master_master over the MASTER groups wired by sync_comms_for_master, then both updates.
Oracle: the FIRST loss of each replica depends only on its initial weights and its half of the batch, so it must equal
a single-process PyTorch fp32 model of the same network (forward through conv_spatial + halo mailboxes + SP->LP
junction on the mirrored rank line); after that the two replicas' gradients are mixed by the MASTER allreduce, so the
later losses are checked to be finite and the run to complete on every rank with libspconv kernels launched."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, SPLIT, IMG, BATCH, STEPS, WIDTH = 2, 3, 64, 2, 3, 8
BALANCE = [5, 2, 2]
SEEDS = (11, 22)


def _layers(conv, pool, seed):
    torch.manual_seed(seed)
    return [conv(3, WIDTH, 3, 1), nn.ReLU(), conv(WIDTH, WIDTH, 3, 2), nn.ReLU(), pool(),
            nn.Conv2d(WIDTH, 4, 3, padding=1), nn.ReLU(),
            nn.Flatten(), nn.Linear(4 * (IMG // 2) ** 2, 10)]


def _batch(step):
    g = torch.Generator().manual_seed(700 + step)
    return torch.randn(2 * BATCH, 3, IMG, IMG, generator=g), torch.randint(0, 10, (2 * BATCH,), generator=g)


def _sequential_first_loss(seed, half):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = nn.Sequential(*_layers(lambda ci, co, k, s: nn.Conv2d(ci, co, k, stride=s, padding=k // 2),
                               lambda: nn.AvgPool2d(3, stride=1, padding=1), seed)).cuda()
    crit = nn.CrossEntropyLoss()
    x, y = _batch(0)
    x, y = x[half * BATCH:(half + 1) * BATCH], y[half * BATCH:(half + 1) * BATCH]
    with torch.no_grad():
        return float(crit(m(x.cuda()), y.cuda()))


def _worker(rank, port, ngpu, q):
    import sys
    sys.path.insert(0, ROOT)
    world = P + SPLIT - 1
    multi = ngpu >= world
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank if multi else 0), SPCONV_DIST_BACKEND="nccl" if multi else "gloo",
                      SPCONV_ARENA_MB="64")
    from mpi4dl_b200.torchgems import comm as gems_comm
    from mpi4dl_b200.torchgems.mp_pipeline import model_generator
    from mpi4dl_b200.torchgems.spatial import Pool, conv_spatial
    from mpi4dl_b200.torchgems.train_spatial import get_shapes_spatial, split_input
    from mpi4dl_b200.torchgems.train_spatial_master import train_spatial_model_master
    gems_comm.initialize_cuda()
    c1 = gems_comm.MPIComm(split_size=SPLIT, ENABLE_MASTER=False, ENABLE_SPATIAL=True, num_spatial_parts=P, spatial_size=1)
    c2 = gems_comm.MPIComm(split_size=SPLIT, ENABLE_MASTER=True, ENABLE_SPATIAL=True, num_spatial_parts=P, spatial_size=1,
                           DISABLE_INIT=True)
    gems_comm.sync_comms_for_master(c1, c2)
    sync = gems_comm.SyncAllreduce(c1)
    full = [(BATCH, WIDTH, IMG // 2, IMG // 2), (BATCH, 4, IMG // 2, IMG // 2), (BATCH, 10)]
    shapes = get_shapes_spatial(full, "vertical", 1, [P], 1)
    gens = []
    for seed, c in zip(SEEDS, (c1, c2)):
        sp = dict(local_rank=c.local_rank, spatial_size=1, num_spatial_parts=P, slice_method="vertical")
        model = nn.Sequential(*_layers(
            lambda ci, co, k, s: conv_spatial(in_channels=ci, out_channels=co, kernel_size=k, stride=s, padding=k // 2, **sp),
            lambda: Pool(operation="AvgPool2d", kernel_size=3, stride=1, padding=1, **sp), seed))
        g = model_generator(model=model, split_size=SPLIT, input_size=(BATCH, 3, IMG, IMG), balance=BALANCE, shape_list=shapes)
        g.ready_model(split_rank=c.split_rank)
        gens.append(g)
    tm = train_spatial_model_master(gens[0], gens[1], BATCH, 1, P, "vertical", c1, c2, LOCAL_DP_LP=1, parts=1, ASYNC=True,
                                    replications=1)
    tm.train_model1.optimizer = torch.optim.SGD(gens[0].models.parameters(), lr=0.005, momentum=0.9)
    tm.train_model2.optimizer = torch.optim.SGD(gens[1].models.parameters(), lr=0.005, momentum=0.9)
    losses = []
    for step in range(STEPS):
        x, y = _batch(step)
        if c1.local_rank < P:
            x = split_input(x, IMG, "vertical", c1.local_rank, [P])
        elif c2.local_rank < P:
            x = split_input(x, IMG, "vertical", c2.local_rank, [P])
        loss, _ = tm.run_step(x, y)
        sync.apply_allreduce_master_master(gens[0], gens[1], c1, c2)
        tm.train_model1.update()
        tm.train_model2.update()
        losses.append(float(loss))
    from mpi4dl_b200 import _lib
    q.put((rank, losses, int(_lib.lib().spc_launch_count(0))))
    dist.barrier()
    dist.destroy_process_group()


def test_gems_master_with_spatial_parallelism_on_gpu():
    want1, want2 = _sequential_first_loss(SEEDS[0], 0), _sequential_first_loss(SEEDS[1], 1)
    world = P + SPLIT - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 29960, torch.cuda.device_count(), q)) for r in range(world)]
    for p in ps:
        p.start()
    got = {}
    import queue
    import time
    deadline = time.time() + 240
    while len(got) < world and time.time() < deadline:
        try:
            r, losses, launches = q.get(timeout=1)
            got[r] = (losses, launches)
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in ps):
                break
    ok = len(got) == world
    for p in ps:
        p.join(30 if ok else 1)
        if p.is_alive():
            p.kill()
    assert ok, "worker exit codes: %s" % [p.exitcode for p in ps]
    assert all(got[r][1] > 0 for r in range(world)), "every rank hosts a tile of one replica: libspconv kernels must have run"
    # replica 1's tail is world rank 3, replica 2's (mirrored line) is world rank 0; run_step returns the tail's loss
    import math
    assert got[world - 1][0][0] == pytest.approx(want1, rel=2e-4, abs=2e-4)
    assert got[0][0][0] == pytest.approx(want2, rel=2e-4, abs=2e-4)
    assert all(math.isfinite(v) for r in (0, world - 1) for v in got[r][0])
