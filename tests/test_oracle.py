"""CPU: the oracle restatement reproduces the reference's own outputs.

(1) against tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference);
(2) live against /root/reference when it is present (build container only).
"""
import numpy as np
import pytest
import torch

from oracle import gpzoo_oracle as O
from oracle import ref_loader
from tests.helpers import GOLDEN, load_golden, oracle_params, relerr

TOL = 1e-10   # fp64


def _check(out, grads, gout, ggrad, keys=("elbo", "ll", "kl", "mean", "var")):
    for k in keys:
        assert relerr(out[k], gout[k]) < TOL, k
    for k, v in ggrad.items():
        assert relerr(grads[k], v) < 1e-9, k


@pytest.mark.parametrize("name", ["nsf_svgp_box", "nsf_svgp_slideseq", "nsf_svgp_1d", "nsf_mggp"])
def test_nsf_svgp_golden(name):
    inp, gout, ggrad = load_golden(name)
    p = oracle_params(inp)
    out, grads = O.value_and_grads(
        lambda: O.nsf_svgp_terms(p, inp["X"], inp["y"], inp["eps"], groupsX=inp.get("groupsX")), p.leaves())
    _check(out, grads, gout, ggrad)
    assert relerr(out["Lc"], gout["Lc"]) < TOL and relerr(out["Lu"], gout["Lu"]) < TOL


def test_nsf_svgp_batched_golden():
    inp, gout, ggrad = load_golden("nsf_svgp_box_batched")
    p = oracle_params(inp)
    idx = inp["idx"]
    out, grads = O.value_and_grads(
        lambda: O.nsf_svgp_terms(p, inp["X"], inp["y"], inp["eps"][:, :, idx], idx=idx, with_lgamma=False), p.leaves())
    _check(out, grads, gout, ggrad)
    assert (grads["V"] != 0).sum() == len(idx)


def test_svgp_gaussian_golden():
    inp, gout, ggrad = load_golden("svgp_gaussian")
    leaves = {k: inp[k].clone() for k in ("Z", "sigma", "lengthscale", "mu", "Lu_raw", "noise")}
    out, grads = O.value_and_grads(
        lambda: O.svgp_gaussian_terms(leaves["Z"], leaves["sigma"], leaves["lengthscale"], leaves["mu"],
                                      leaves["Lu_raw"], leaves["noise"], inp["X"], inp["y"], inp["eps"], inp["jitter"]),
        leaves)
    _check(out, grads, gout, ggrad)


def test_vnngp_golden():
    inp, gout, ggrad = load_golden("nsf_vnngp")
    p = oracle_params(inp)
    out, grads = O.value_and_grads(lambda: O.vnngp_terms(p, inp["X"], inp["y"], inp["eps"], inp["K"]), p.leaves())
    assert torch.equal(out["nn"], gout["nn"])            # neighbour indices: bit-exact
    _check(out, grads, gout, ggrad)


def test_hybrid_golden():
    inp, gout, ggrad = load_golden("nsf_hybrid")
    p = oracle_params(inp)
    idx = inp["idx"]
    extra = {k: inp[k].clone() for k in ("Wcf", "cf_mean", "cf_scale")}
    leaves = dict(p.leaves(), **extra)
    out, grads = O.value_and_grads(
        lambda: O.hybrid_terms(p, extra["Wcf"], extra["cf_mean"], extra["cf_scale"], inp["X"], inp["y"],
                               inp["eps"][:, :, idx], inp["eps2"][:, :, idx], idx=idx), leaves)
    _check(out, grads, gout, ggrad)
    assert relerr(out["kl2"].sum(), gout["kl2"]) < TOL


def test_kernel_goldens():
    z = {k: torch.from_numpy(v) for k, v in np.load(GOLDEN + "/kernels.npz").items()}
    X, Z, gX, gZ = z["X"], z["Z"], z["gX"], z["gZ"]
    t = lambda v: torch.tensor(v).double()      # the reference's ctor rounds params to fp32 first
    assert relerr(O.rbf(X, Z, t(1.3), t(0.7)), z["rbf"]) < TOL
    assert relerr(O.rbf_diag(X, t(1.3)), z["rbf_diag"]) < TOL
    sig = (1.1 * torch.ones(3, 1, 1)).double()
    assert relerr(O.nsf_rbf(X, Z, sig, z["nsf_rbf_ls"]), z["nsf_rbf"]) < TOL
    emb = O.embed_distance_matrix(z["gd"].float()).double()
    assert relerr(emb @ emb.t(), z["mggp_embedding"] @ z["mggp_embedding"].t()) < 1e-6   # eigvec sign-free
    emb = z["mggp_embedding"]
    assert relerr(O.mggp_rbf(X, Z, gX, gZ, t(1.2), t(0.8), t(1.7), emb), z["mggp_rbf"]) < TOL
    one = torch.ones(2, 1, 1)
    assert relerr(O.mggp_nsf_rbf(X, Z, gX, gZ, (1.2 * one).double(), (0.8 * one).double(), (1.3 * one).double(), emb), z["mggp_nsf_rbf"]) < TOL
    assert relerr(O.matern32(X, Z, t(1.2), t(0.8)), z["matern32"]) < TOL
    assert relerr(O.whitened_kl(z["mz"], z["Lz"]), z["whitened_kl"]) < TOL
    Kzx = O.nsf_rbf(Z, X, sig, z["nsf_rbf_ls"])
    Kzz = O.nsf_rbf(Z, Z, sig, z["nsf_rbf_ls"])
    mean, var, _ = O.wsvgp(O.rbf_diag(X, sig), Kzx, Kzz, z["wsvgp_mu"], z["wsvgp_Lu"], 1e-2)
    assert relerr(mean, z["wsvgp_mean"]) < TOL and relerr(var, z["wsvgp_var"]) < TOL
    # default group distances (ones - eye): r^2 = 0 within a group, ~1.000002 across
    d_emb = z["mggp_default_embedding"]
    r2 = O.squared_dist(d_emb, d_emb)
    assert abs(r2[0, 1].item() - 1.000002) < 1e-5 and abs(r2[0, 0].item()) < 1e-6


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (GPU box)")
def test_oracle_matches_live_reference_fp32_and_fp64():
    from gpzoo_b200 import synthetic
    from oracle import ref_runner
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 2e-4)):
        prob = synthetic.nsf_problem(N=200, M=49, L=3, G=20, E=2, seed=3, coord_scale=2.0, jitter=1e-2, dtype=dtype)
        rout, rgrad = ref_runner.run_nsf_svgp(prob)
        p = oracle_params(prob, dtype)
        out, grads = O.value_and_grads(lambda: O.nsf_svgp_terms(p, prob["X"], prob["y"], prob["eps"]), p.leaves())
        for k in ("elbo", "ll", "kl", "mean", "var"):
            assert relerr(out[k], rout[k]) < tol, (k, dtype)
        for k, v in rgrad.items():
            assert relerr(grads[k], v) < 10 * tol, (k, dtype)


def test_init_pipeline_golden():
    """The deterministic part of the initialisation pipeline (utilities.py:237-313, 38-44, 71-84) against the reference's outputs."""
    z = np.load(GOLDEN + "/init_pipeline.npz")
    t = lambda k: torch.from_numpy(z[k])
    for name, shrink in (("mu_kl", 0.2), ("mu_fro", 0.3), ("cd", 0.2)):
        Fl, W = O.regularized_nmf_post(t(name + "_eF"), t(name + "_H").t(), 4, sz=t("sz"), shrinkage=shrink)
        assert relerr(Fl, t(name + "_F")) < TOL and relerr(W, t(name + "_W")) < TOL, name
    Fl, W = O.regularized_nmf_post(t("cd_eF"), t("cd_H").t(), 4, shrinkage=0.25)
    assert relerr(Fl, t("post_F")) < TOL and relerr(W, t("post_W")) < TOL
    assert relerr(O.init_softplus(t("softplus_in")), t("softplus_out")) < TOL
    assert relerr(O.rescale_spatial_coords(t("coords_in")), t("coords_out")) < TOL
    for l, ref in zip((1.1, 4, 10), z["lnormal"]):
        assert np.allclose(O.lnormal_approx_dirichlet(l), ref, rtol=1e-14, atol=1e-14)


@pytest.mark.parametrize("solver,beta_loss", [("mu", "kullback-leibler"), ("mu", "frobenius"), ("cd", "frobenius")])
def test_nmf_oracle_vs_sklearn(solver, beta_loss):
    """The numpy restatement of sklearn's NMF solvers (the third-party algorithm behind utilities.py:253-299) against sklearn itself,
    from the same start, including the iteration at which the stopping rule fires."""
    import warnings
    from sklearn.decomposition import NMF
    from oracle import nmf_oracle as NO
    Y = np.load(GOLDEN + "/init_pipeline.npz")["Y"]
    rng = np.random.RandomState(7)
    W0, H0 = np.abs(rng.standard_normal((Y.shape[0], 4))) + 0.1, np.abs(rng.standard_normal((4, Y.shape[1]))) + 0.1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = NMF(4, init="custom", solver=solver, beta_loss=beta_loss, max_iter=300, tol=1e-4)
        Wr = m.fit_transform(Y.copy(), W=W0.copy(), H=H0.copy())
    if solver == "mu":
        W, H, n = NO.nmf_mu(Y, W0, H0, 1 if beta_loss == "kullback-leibler" else 2, max_iter=300, tol=1e-4)
    else:
        W, H, n = NO.nmf_cd(Y, W0, H0, max_iter=300, tol=1e-4)
    assert n == m.n_iter_
    assert relerr(W, Wr) < 1e-9 and relerr(H, m.components_) < 1e-9
