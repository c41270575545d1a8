"""Structure fixtures of the UNMODIFIED reference's D2 ResNet builder (src/models/resnet_spatial_d2.py) on CPU
-> tests/golden/model_d2_golden.json.  Build container only (imports /root/reference under tools/ref_shim.py).
Per configuration: state-dict signature, parameter count, the ordered kinds of conv / halo modules, the
(name, halo_len) of every inserted halo_exchange_layer and the balance the builder returns."""
import json
import os
import sys
import warnings

import torch.distributed as dist

warnings.simplefilter("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "model_d2_golden.json")


def main():
    sys.path.insert(0, HERE)
    import ref_shim
    ref_shim.install()
    from gen_model_golden import kinds, sig
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29778")
    dist.init_process_group("gloo", rank=0, world_size=1)
    from models import resnet_spatial_d2
    out = []
    for depth in (29, 56, 101):
        nb = (depth - 2) // 9
        for fused in sorted({1, 2, min(3, nb), nb}):
            for kw in (dict(mp_size=2, balance=None), dict(mp_size=3, balance=None),
                       dict(mp_size=3, balance=[nb + 2, nb, 3 * nb + 2 - 2 * nb - 2])):
                m, bal = resnet_spatial_d2.get_resnet_v2((2, 3, 64, 64), depth, local_rank=0, spatial_size=1, num_spatial_parts=4,
                                                         slice_method="square", fused_layers=fused,
                                                         balance=list(kw["balance"]) if kw["balance"] else None, mp_size=kw["mp_size"])
                kh, nconv, _ = kinds(m)
                halos = [(n, x.halo_len) for n, x in m.named_children() if type(x).__name__ == "halo_exchange_layer"]
                out.append(dict(depth=depth, fused_layers=fused, kw=kw, state_sig=sig(m), params=sum(p.numel() for p in m.parameters()),
                                kinds_sig=kh, spatial_convs=nconv, halos=halos, balance=list(bal), children=[n for n, _ in m.named_children()]))
    json.dump({"source": "tools/gen_model_d2_golden.py on unmodified /root/reference/src/models/resnet_spatial_d2.py (CPU)",
               "resnet_d2": out}, open(OUT, "w"), indent=1)
    print(len(out), "configurations")


if __name__ == "__main__":
    main()
