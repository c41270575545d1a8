"""CPU-only (gloo): mpi4dl_b200.torchgems.comm against attribute tables generated from the
UNMODIFIED reference's MPIComm (tools/gen_comm_golden.py -> tests/golden/comm_golden.json), plus
the numerics of SyncAllreduce.apply_allreduce (reference comm.py:506-514: sum over the spatial
group, divided by num_spatial_parts)."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "comm_golden.json")))["configs"]


def _grp(g):
    return None if g is None else sorted(int(r) for r in dist.get_process_group_ranks(g))


def _dump(c):
    d = {}
    for a in ("mp_size", "rank", "size", "local_rank", "split_rank", "total_spatial_processes", "split_size"):
        if hasattr(c, a):
            v = getattr(c, a)
            d[a] = int(v) if v is not None else None
    for a in ("spatial_allreduce_grp", "allreduce_grp", "SP_LP_group", "LOCAL_DP_MP_Comm", "first_spatial_allreduce_grp",
              "second_spatial_allreduce_grp", "first_LP_master_group", "second_LP_master_group", "allreduce_grp_master"):
        if hasattr(c, a):
            d[a] = _grp(getattr(c, a))
    if getattr(c, "LP_SP_Groups", None) is not None:
        d["LP_SP_Groups"] = [_grp(g) for g in c.LP_SP_Groups]
    return d


def _worker(rank, world, kw, master, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SPCONV_DIST_BACKEND="gloo")
    from mpi4dl_b200.torchgems import comm
    c1 = comm.MPIComm(**kw)
    out = {}
    if master:
        c2 = comm.MPIComm(ENABLE_MASTER=True, DISABLE_INIT=True, **kw)
        comm.sync_comms_for_master(c1, c2)
        out["comm2"] = _dump(c2)
    out["comm1"] = _dump(c1)
    s = comm.SyncAllreduce(c1)
    out["divide_bs"] = float(s.divide_bs)
    # numerics: every rank holds grad = rank+1 on a tiny model; spatial ranks allreduce
    if kw.get("ENABLE_SPATIAL") and not master and c1.spatial_allreduce_grp is not None \
            and c1.local_rank < c1.total_spatial_processes and not isinstance(kw["num_spatial_parts"], list):
        m = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.Linear(2, 1))
        for p in m.parameters():
            p.grad = torch.full_like(p, float(rank + 1))

        class G:
            models = m
        s.apply_allreduce(G, c1.spatial_allreduce_grp)
        P = kw["num_spatial_parts"]
        expect = sum(r + 1 for r in range(P)) / P
        out["allreduce_ok"] = all(torch.allclose(p.grad, torch.full_like(p, expect)) for p in m.parameters())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, kw, master, port):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    ps = [ctx.Process(target=_worker, args=(r, world, kw, master, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get() for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    return [got[r] for r in range(world)]


@pytest.mark.parametrize("idx,name", list(enumerate(sorted(GOLD))))
def test_mpicomm_matches_reference_tables(idx, name):
    cfg = GOLD[name]
    mine = _run(cfg["world"], cfg["kw"], cfg["master"], 29950 + idx)
    for r, (m, g) in enumerate(zip(mine, cfg["ranks"])):
        for which in ("comm1", "comm2"):
            if which not in g:
                continue
            for k, v in g[which].items():
                if isinstance(v, list) and v and isinstance(v[0], int) and r not in v:
                    continue   # handle of a group this rank is not a member of: membership is undefined
                if k == "LP_SP_Groups":
                    v = [x for x in v if r in x]
                    got = [x for x in (m[which].get(k) or []) if x and r in x]
                    assert got == v, (name, r, which, k, got, v)
                    continue
                assert m[which].get(k) == v, (name, r, which, k, m[which].get(k), v)
        assert m["divide_bs"] == g["divide_bs"]
        if "allreduce_ok" in m:
            assert m["allreduce_ok"], (name, r)


def test_data_parallel_groups_by_formula():
    """world = 2 * mp_size: DP allreduce groups are the ranks with equal local_rank
    (comm.py:161-168).  (The reference cannot build this under gloo: it creates a different group
    list on every rank, which only MPI tolerates.)"""
    res = _run(4, dict(split_size=2), False, 29990)
    for r, m in enumerate(res):
        assert m["comm1"]["mp_size"] == 2 and m["comm1"]["local_rank"] == r % 2
        assert m["comm1"]["allreduce_grp"] == [r % 2, r % 2 + 2]
        assert m["divide_bs"] == 2.0
