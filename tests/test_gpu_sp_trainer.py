"""-m gpu: the SP+LP training step end to end through the caller stack -- MPIComm, model_generator,
train_model_spatial (tiles -> join rank -> tail), SyncAllreduce -- with the spatial stage on
conv_spatial / Pool (libspconv kernels + halo exchange) on P=2 tiles, checked against a
single-process PyTorch fp32 model of the same network trained with the same rule (spatial-stage
gradients are SUM over tiles / P, the reference's convention, comm.py:440-458).

World = 4 processes (2 tiles + join + tail).  With >= 4 GPUs each gets its own device over NCCL;
on fewer GPUs they share cuda:0 and torch.distributed runs on gloo (host-staged sends) -- the
kernels, the halo mailboxes and the trainer logic are the same."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, SPLIT, IMG, BATCH, STEPS = 2, 3, 128, 2, 3
BALANCE = [5, 2, 2]


def _layers(conv, pool, width):
    """Same construction order for both builds -> same default init under one seed."""
    torch.manual_seed(99)
    return [conv(3, width, 3, 1), nn.ReLU(), conv(width, width, 3, 2), nn.ReLU(), pool(),   # spatial stage
            nn.Conv2d(width, 4, 3, padding=1), nn.ReLU(),                                 # join rank
            nn.Flatten(), nn.Linear(4 * (IMG // 2) ** 2, 10)]                            # tail


def _batch(step):
    g = torch.Generator().manual_seed(500 + step)
    return torch.randn(BATCH, 3, IMG, IMG, generator=g), torch.randint(0, 10, (BATCH,), generator=g)


def _sequential_losses(width):
    torch.backends.cudnn.allow_tf32 = False          # true-fp32 baseline (cuDNN defaults to TF32 convolutions)
    torch.backends.cuda.matmul.allow_tf32 = False
    m = nn.Sequential(*_layers(lambda ci, co, k, s: nn.Conv2d(ci, co, k, stride=s, padding=k // 2),
                               lambda: nn.AvgPool2d(3, stride=1, padding=1), width)).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=0.005, momentum=0.9)
    crit = nn.CrossEntropyLoss()
    n_spatial = sum(1 for _ in nn.Sequential(*list(m)[:BALANCE[0]]).parameters())
    losses = []
    for step in range(STEPS):
        x, y = _batch(step)
        loss = crit(m(x.cuda()), y.cuda())
        loss.backward()
        for i, p in enumerate(m.parameters()):
            if i < n_spatial:
                p.grad.div_(P)
        opt.step()
        opt.zero_grad()
        losses.append(loss.item())
    return losses


def _worker(rank, method, width, dtype, port, ngpu, q):
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
    from mpi4dl_b200.torchgems.train_spatial import get_shapes_spatial, split_input, train_model_spatial
    gems_comm.initialize_cuda()
    mpi_comm = gems_comm.MPIComm(split_size=SPLIT, ENABLE_MASTER=False, ENABLE_SPATIAL=True, num_spatial_parts=P, spatial_size=1)
    sync = gems_comm.SyncAllreduce(mpi_comm)
    local_rank, split_rank = mpi_comm.rank, mpi_comm.split_rank
    sp = dict(local_rank=local_rank % P, spatial_size=1, num_spatial_parts=P, slice_method=method)
    model = nn.Sequential(*_layers(
        lambda ci, co, k, s: conv_spatial(in_channels=ci, out_channels=co, kernel_size=k, stride=s, padding=k // 2, **sp),
        lambda: Pool(operation="AvgPool2d", kernel_size=3, stride=1, padding=1, **sp), width)).to(getattr(torch, dtype))
    # full-image stage shapes, then tiled
    full = [(BATCH, width, IMG // 2, IMG // 2), (BATCH, 4, IMG // 2, IMG // 2), (BATCH, 10)]
    shapes = get_shapes_spatial(full, method, 1, [P], 1)
    gen = model_generator(model=model, split_size=SPLIT, input_size=(BATCH, 3, IMG, IMG), balance=BALANCE, shape_list=shapes)
    gen.ready_model(split_rank=split_rank)
    opt = torch.optim.SGD(gen.models.parameters(), lr=0.005, momentum=0.9)
    tm = train_model_spatial(gen, local_rank, BATCH, epochs=1, spatial_size=1, num_spatial_parts=P, optimizer=opt,
                             parts=1, slice_method=method, mpi_comm=mpi_comm)
    sync.sync_model_spatial(gen)
    losses = []
    for step in range(STEPS):
        x, y = _batch(step)
        if local_rank < P:
            x = split_input(x, IMG, method, local_rank, [P])
        loss, _ = tm.run_step(x, y)
        if local_rank < P:
            sync.apply_allreduce(gen, mpi_comm.spatial_allreduce_grp)
        tm.update()
        losses.append(float(loss))
    from mpi4dl_b200 import _lib
    q.put((local_rank, losses, int(_lib.lib().spc_launch_count(0))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("idx,method,width,dtype,tol", [(0, "vertical", 8, "float32", 2e-4), (1, "horizontal", 8, "float32", 2e-4),
                                                         (2, "vertical", 64, "bfloat16", 5e-2)])
def test_sp_lp_training_matches_single_process(idx, method, width, dtype, tol):
    want = _sequential_losses(width)
    world = P + SPLIT - 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, method, width, dtype, 29940 + idx, torch.cuda.device_count(), q))
          for r in range(world)]
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
    assert got[0][1] > 0 and got[1][1] > 0, "tile ranks did not run libspconv kernels"
    assert got[world - 1][0] == pytest.approx(want, rel=tol, abs=tol)
