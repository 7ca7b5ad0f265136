"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only from
/root/reference) on small seeded problems.  Run once in the build container:

    python -m oracle.gen_golden

The reference ships no golden vectors of its own, so these files — inputs,
ELBO pieces and gradients produced by the reference's own modules in fp64 — are what pins the
oracle (tests/test_oracle.py) and, through it and directly, the CUDA path (tests/test_*_gpu.py).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from gpzoo_b200 import synthetic  # noqa: E402
from oracle import ref_runner     # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(d, prefix):
    out = {}
    for k, v in d.items():
        if torch.is_tensor(v):
            out[prefix + k] = v.detach().cpu().numpy()
        else:
            out[prefix + k] = np.asarray(v)
    return out


def save(name, prob, out, grads, **extra):
    os.makedirs(OUT, exist_ok=True)
    blob = {}
    blob.update(_np(prob, "in_"))
    blob.update(_np(out, "out_"))
    blob.update(_np(grads, "grad_"))
    blob.update(_np(extra, "in_"))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
    print(name, "elbo=%.12g" % float(out["elbo"]) if "elbo" in out else "")


CASES = {
    # name: kwargs of synthetic.nsf_problem
    "nsf_svgp_box": dict(N=96, M=25, L=3, G=12, E=2, seed=11, coord_scale=2.0, jitter=1e-2),
    "nsf_svgp_slideseq": dict(N=128, M=36, L=2, G=10, E=1, seed=12, coord_scale=100.0, lengthscale=25.0, jitter=1e-1),
    "nsf_svgp_1d": dict(N=64, M=12, L=2, G=6, E=2, D=1, seed=13, coord_scale=3.0, jitter=1e-2),
    "nsf_mggp": dict(N=96, M=24, L=2, G=8, E=2, seed=14, coord_scale=2.0, jitter=1e-2, n_groups=3),
}


def main():
    torch.manual_seed(0)
    ref = load_reference()
    for name, kw in CASES.items():
        prob = synthetic.nsf_problem(**kw)
        out, grads = ref_runner.run_nsf_svgp(prob)
        save(name, prob, out, grads)
    # minibatched (forward_batched) variant of the first case, y*log(rate)-rate ELBO form
    prob = synthetic.nsf_problem(**CASES["nsf_svgp_box"])
    idx = torch.randperm(96, generator=torch.Generator().manual_seed(5))[:40]
    out, grads = ref_runner.run_nsf_svgp(prob, idx=idx, with_lgamma=False)
    save("nsf_svgp_box_batched", prob, out, grads, idx=idx)

    prob = synthetic.regression_problem(N=80, M=16, E=3, seed=21)
    out, grads = ref_runner.run_svgp_gaussian(prob)
    save("svgp_gaussian", prob, out, grads)

    prob = synthetic.nsf_problem(N=64, M=30, L=2, G=8, E=2, seed=31, coord_scale=2.0, jitter=1e-2, lu_scale=0.02)
    out, grads = ref_runner.run_vnngp(prob, K=4)
    save("nsf_vnngp", prob, out, grads, K=4)

    prob = synthetic.nsf_problem(N=64, M=16, L=2, G=8, E=2, seed=41, coord_scale=2.0, jitter=1e-2, lengthscale=0.8)
    g = torch.Generator().manual_seed(42)
    extra = dict(Wcf=torch.rand(8, 3, generator=g, dtype=torch.float64),
                 cf_mean=0.3 * torch.randn(3, 64, generator=g, dtype=torch.float64),
                 cf_scale=torch.rand(3, 64, generator=g, dtype=torch.float64),
                 eps2=torch.randn(2, 3, 64, generator=g, dtype=torch.float64))
    idx = torch.randperm(64, generator=g)[:40]
    out, grads = ref_runner.run_hybrid(prob, extra, idx)
    save("nsf_hybrid", prob, out, grads, idx=idx, **extra)

    # kernel-level goldens (kernels.py:106-228, 6-30) and small helpers (utilities.py:27-36; gp.py:260-306)
    g = torch.Generator().manual_seed(51)
    X = 2 * torch.rand(40, 2, generator=g, dtype=torch.float64) - 1
    Z = 2 * torch.rand(17, 2, generator=g, dtype=torch.float64) - 1
    gX = torch.randint(0, 4, (40,), generator=g)
    gZ = torch.randint(0, 4, (17,), generator=g)
    gd = torch.rand(4, 4, generator=g, dtype=torch.float64) + 0.5
    gd = 0.5 * (gd + gd.t())
    gd.fill_diagonal_(0)
    k = {}
    with torch.no_grad():
        r = ref.kernels.RBF(sigma=1.3, lengthscale=0.7).double()
        k["rbf"] = r(X, Z)
        k["rbf_diag"] = r(X, X, diag=True)
        nr = ref.kernels.NSF_RBF(sigma=1.1, lengthscale=0.9, L=3).double()
        nr.lengthscale.mul_(torch.tensor([1.0, 1.5, 2.0], dtype=torch.float64).reshape(3, 1, 1))
        k["nsf_rbf"] = nr(X, Z)
        k["nsf_rbf_ls"] = nr.lengthscale.detach()
        mr = ref.kernels.MGGP_RBF(sigma=1.2, lengthscale=0.8, group_diff_param=1.7, n_groups=4).double()
        k["mggp_default_embedding"] = mr.embedding.double()
        k["mggp_rbf_default"] = ref.kernels.MGGP_RBF.forward(mr, X.float(), Z.float(), gX, gZ).double()
        mr.embedding = ref.utilities._embed_distance_matrix(gd.float()).double()   # reference builds it in fp32 (utilities.py:464)
        k["mggp_embedding"] = mr.embedding
        k["mggp_rbf"] = mr(X, Z, gX, gZ)
        mn = ref.kernels.MGGP_NSF_RBF(sigma=1.2, lengthscale=0.8, group_diff_param=1.3, n_groups=4, L=2).double()
        mn.embedding = torch.nn.Parameter(ref.utilities._embed_distance_matrix(gd.float()).double(), requires_grad=False)
        k["mggp_nsf_rbf"] = mn(X, Z, gX, gZ)
        mt = ref.kernels.batched_Matern32(sigma=1.2, lengthscale=0.8).double()
        k["matern32"] = mt(X, Z)
        mz = torch.randn(17, generator=g, dtype=torch.float64)
        Lz = torch.tril(torch.randn(17, 17, generator=g, dtype=torch.float64)) * 0.1 + torch.eye(17, dtype=torch.float64)
        k["whitened_kl"] = ref.utilities.whitened_KL(mz, Lz)
        # WSVGP (gp.py:235-306), L-batched
        w = ref.gp.WSVGP(nr, dim=2, M=17, jitter=1e-2).double()
        w.Z = torch.nn.Parameter(Z.clone())
        w.mu = torch.nn.Parameter(0.5 * torch.randn(3, 17, generator=g, dtype=torch.float64))
        w.Lu = torch.nn.Parameter(0.05 * torch.randn(3, 17, 17, generator=g, dtype=torch.float64))
        qF, qZ, _ = w(X)
        k["wsvgp_mean"], k["wsvgp_var"] = qF.mean, qF.scale ** 2
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), X=X.numpy(), Z=Z.numpy(), gX=gX.numpy(), gZ=gZ.numpy(),
                        gd=gd.numpy(), mz=mz.numpy(), Lz=Lz.numpy(), wsvgp_mu=w.mu.detach().numpy(),
                        wsvgp_Lu=w.Lu.detach().numpy(), **{kk: vv.detach().numpy() for kk, vv in k.items()})
    print("kernels.npz written")


if __name__ == "__main__":
    main()
