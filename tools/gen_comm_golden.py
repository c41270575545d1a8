"""Attribute tables of the UNMODIFIED reference's torchgems.comm.MPIComm, generated on CPU/gloo
(tools/ref_shim.py) for several launch configurations -> tests/golden/comm_golden.json.
Run in the build container only:  python tools/gen_comm_golden.py"""
import json
import os
import sys

import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "comm_golden.json")

CONFIGS = [
    dict(name="sp2_lp2", world=3, kw=dict(split_size=2, ENABLE_SPATIAL=True, num_spatial_parts=2, spatial_size=1)),
    dict(name="sp4_lp3", world=6, kw=dict(split_size=3, ENABLE_SPATIAL=True, num_spatial_parts=4, spatial_size=1)),
    dict(name="lp2", world=2, kw=dict(split_size=2)),
    dict(name="sp22_lp3", world=5, kw=dict(split_size=3, ENABLE_SPATIAL=True, num_spatial_parts=[2, 2], spatial_size=2)),
    dict(name="sp2_lp2_localdp2", world=4, kw=dict(split_size=2, ENABLE_SPATIAL=True, num_spatial_parts=2, spatial_size=1, LOCAL_DP_LP=2)),
    dict(name="master_sp2_lp3", world=4, master=True, kw=dict(split_size=3, ENABLE_SPATIAL=True, num_spatial_parts=2, spatial_size=1)),
]


def grp(g):
    return None if g is None else sorted(int(r) for r in dist.get_process_group_ranks(g))


def dump(c):
    d = {}
    for a in ("mp_size", "rank", "size", "local_rank", "split_rank", "total_spatial_processes", "split_size"):
        if hasattr(c, a):
            v = getattr(c, a)
            d[a] = int(v) if v is not None else None
    for a in ("spatial_allreduce_grp", "allreduce_grp", "SP_LP_group", "LOCAL_DP_MP_Comm", "first_spatial_allreduce_grp",
              "second_spatial_allreduce_grp", "first_LP_master_group", "second_LP_master_group", "allreduce_grp_master"):
        if hasattr(c, a):
            d[a] = grp(getattr(c, a))
    if getattr(c, "LP_SP_Groups", None) is not None:
        d["LP_SP_Groups"] = [grp(g) for g in c.LP_SP_Groups]
    return d


def worker(rank, cfg, port, q):
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)   # the reference's per-rank new_group lists can hang on gloo
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.install()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(cfg["world"]))
    from torchgems import comm as ref
    c1 = ref.MPIComm(**cfg["kw"])
    out = {"comm1": None}
    if cfg.get("master"):
        c2 = ref.MPIComm(ENABLE_MASTER=True, DISABLE_INIT=True, **cfg["kw"])
        ref.sync_comms_for_master(c1, c2)
        out["comm2"] = dump(c2)
    out["comm1"] = dump(c1)
    s = ref.SyncAllreduce(c1)
    out["divide_bs"] = float(s.divide_bs)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ctx = mp.get_context("spawn")
    res = {}
    port = 29900
    for cfg in CONFIGS:
        port += 1
        q = ctx.SimpleQueue()
        ps = [ctx.Process(target=worker, args=(r, cfg, port, q)) for r in range(cfg["world"])]
        for p in ps:
            p.start()
        got = dict(q.get() for _ in ps)
        for p in ps:
            p.join()
        res[cfg["name"]] = {"world": cfg["world"], "kw": cfg["kw"], "master": bool(cfg.get("master")),
                            "ranks": [got[r] for r in range(cfg["world"])]}
        print("done", cfg["name"], flush=True)
    json.dump({"source": "tools/gen_comm_golden.py on unmodified /root/reference src/torchgems/comm.py (gloo)", "configs": res},
              open(OUT, "w"), indent=1, default=lambda o: int(o))
    print("wrote", OUT)


if __name__ == "__main__":
    main()
