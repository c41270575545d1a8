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
    json.dump({"source": "tools/gen_trainer_golden.py on unmodified /root/reference mp_pipeline.py (gloo, CPU)", "cases": res},
              open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
