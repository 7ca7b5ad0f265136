"""Shared test helpers: golden-fixture loading, relative-L2 error, oracle drivers."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, dtype=torch.float64):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    inp, out, grad = {}, {}, {}
    for k in z.files:
        v = torch.from_numpy(np.asarray(z[k]))
        if v.is_floating_point():
            v = v.to(dtype)
        if k.startswith("in_"):
            inp[k[3:]] = v
        elif k.startswith("out_"):
            out[k[4:]] = v
        elif k.startswith("grad_"):
            grad[k[5:]] = v
    if "jitter" in inp:
        inp["jitter"] = float(inp["jitter"])
    if "K" in inp:
        inp["K"] = int(inp["K"])
    return inp, out, grad


def relerr(a, b):
    """Relative L2 error of a against b (b is the truth)."""
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num


def oracle_params(inp, dtype=torch.float64):
    from oracle import gpzoo_oracle as O
    emb = None
    if "group_distances" in inp:
        emb = O.embed_distance_matrix(inp["group_distances"].float()).to(dtype)   # reference builds it in fp32
    f = lambda k: inp[k].to(dtype).clone()
    return O.NSFParams(Z=f("Z"), sigma=f("sigma"), lengthscale=f("lengthscale"), mu=f("mu"), Lu_raw=f("Lu_raw"),
                       W=f("W"), V=f("V"), jitter=inp["jitter"],
                       gdp=f("gdp") if "gdp" in inp else None, embedding=emb,
                       groupsZ=inp.get("groupsZ"))
