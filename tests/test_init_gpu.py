"""GPU: the initialisation pipeline (gpzoo_b200.initialisation) against the reference's outputs
(tests/golden/init_pipeline.npz, written by oracle/gen_golden.py --init from the unmodified reference, which calls sklearn) and
against sklearn itself on the same start."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN + "/init_pipeline.npz")


def test_small_formulas(gold):
    from gpzoo_b200 import utilities as U
    for l, ref in zip((1.1, 4, 10), gold["lnormal"]):
        assert np.allclose(U.lnormal_approx_dirichlet(l), ref, rtol=1e-14, atol=1e-14)
    assert relerr(U.init_softplus(gold["softplus_in"]), gold["softplus_out"]) < 1e-12
    assert relerr(U.rescale_spatial_coords(gold["coords_in"].copy()), gold["coords_out"]) < 1e-12
    X = torch.from_numpy(gold["coords_out"]).float().to(DEV)
    gd = U.build_group_distances(X, torch.from_numpy(gold["groups"]).to(DEV))
    assert relerr(gd, gold["group_distances"]) < 1e-5


def test_regularized_nmf_postprocessing(gold):
    """factors / loadings given: the shrinkage + log-scale normalisation alone (utilities.py:284-299), numpy in -> numpy out."""
    from gpzoo_b200 import utilities as U
    Fl, W = U.regularized_nmf(gold["Y"], 4, sz=1, pseudocount=1e-2, factors=gold["cd_eF"], loadings=gold["cd_H"].T, shrinkage=0.25)
    assert isinstance(Fl, np.ndarray) and relerr(Fl, gold["post_F"]) < 1e-12 and relerr(W, gold["post_W"]) < 1e-12


@pytest.mark.parametrize("name,kw", [
    ("mu_fro", dict(shrinkage=0.3, max_iter=150, solver="mu", init="random", beta_loss="frobenius", random_state=5)),
    ("cd", dict(shrinkage=0.2, max_iter=200, init="random", random_state=1)),
])
def test_regularized_nmf_same_start_as_reference(gold, name, kw):
    """init='random' draws the same numpy stream as sklearn, so the whole call reproduces the reference's (sklearn's) factors."""
    from gpzoo_b200 import initialisation as I
    nkw = {k: v for k, v in kw.items() if k != "shrinkage"}
    W, H, n_iter = I.nmf(gold["Y"], 4, return_n_iter=True, **nkw)
    assert n_iter == int(gold[name + "_n_iter"])
    assert relerr(W, gold[name + "_eF"]) < 1e-7 and relerr(H, gold[name + "_H"]) < 1e-7
    Fl, Wl = I.regularized_nmf(gold["Y"], 4, sz=gold["sz"], **kw)
    assert relerr(Fl, gold[name + "_F"]) < 1e-7 and relerr(Wl, gold[name + "_W"]) < 1e-7


def test_nmf_mu_kl_custom_start_vs_sklearn(gold):
    """Kullback-Leibler multiplicative updates (the notebooks' solver) from a given start: sklearn on the host vs the device."""
    from sklearn.decomposition import NMF
    from gpzoo_b200 import initialisation as I
    rng = np.random.RandomState(0)
    Y = gold["Y"]
    W0, H0 = np.abs(rng.standard_normal((Y.shape[0], 5))) + 0.1, np.abs(rng.standard_normal((5, Y.shape[1]))) + 0.1
    m = NMF(5, init="custom", solver="mu", beta_loss="kullback-leibler", max_iter=120, tol=1e-4)
    Wr = m.fit_transform(Y.copy(), W=W0.copy(), H=H0.copy())
    W, H, n_iter = I.nmf(Y, 5, init="custom", W=W0, H=H0, solver="mu", beta_loss="kullback-leibler", max_iter=120, tol=1e-4,
                         return_n_iter=True)
    assert n_iter == m.n_iter_
    assert relerr(W, Wr) < 1e-8 and relerr(H, m.components_) < 1e-8


def test_nndsvd_inits_vs_sklearn(gold, monkeypatch):
    """NNDSVD starts.  sklearn builds them from a RANDOMISED truncated SVD (accurate to ~3e-3 on the trailing components of this
    matrix); here the SVD is exact.  With sklearn's SVD swapped for an exact one its construction and ours agree to rounding,
    including the random fill of nndsvdar; against stock sklearn the starts agree to the accuracy of its SVD."""
    import sklearn.decomposition._nmf as sk
    from gpzoo_b200 import initialisation as I
    Y = gold["Y"]
    Yd = torch.from_numpy(Y).to(DEV)
    stock = {init: sk._initialize_nmf(Y, 4, init=init, random_state=0) for init in ("nndsvd", "nndsvda", "nndsvdar")}

    def exact_svd(X, k, random_state=None, **kw):
        U, S, Vt = np.linalg.svd(X, full_matrices=False)
        return U[:, :k], S[:k], Vt[:k]
    monkeypatch.setattr(sk, "_randomized_svd", exact_svd)
    for init in ("nndsvd", "nndsvda", "nndsvdar"):
        Wr, Hr = sk._initialize_nmf(Y, 4, init=init, random_state=0)
        W, H = I.initialize_nmf(Yd, 4, init=init, random_state=0)
        assert relerr(W, Wr) < 1e-9 and relerr(H, Hr) < 1e-9, init
        assert relerr(W, stock[init][0]) < 2e-2 and relerr(H, stock[init][1]) < 2e-2, init
    # the notebooks' call (NSF_Hybrid_benchmark.ipynb cell 7): 200 Kullback-Leibler updates from nndsvdar reach the same objective
    # as the reference's (sklearn's) factors
    W, H = I.nmf(Yd, 4, max_iter=200, solver="mu", init="nndsvdar", beta_loss="kullback-leibler", random_state=0)
    ours = I._beta_divergence(Yd, W, H, 1)
    ref = I._beta_divergence(Yd, torch.from_numpy(gold["mu_kl_eF"]).to(DEV), torch.from_numpy(gold["mu_kl_H"]).to(DEV), 1)
    assert abs(ours / ref - 1) < 1e-3
    Fl, Wl = I.regularized_nmf(Y, 4, sz=gold["sz"], shrinkage=0.2, max_iter=200, solver="mu", init="nndsvdar",
                               beta_loss="kullback-leibler", random_state=0)
    assert relerr(Fl, gold["mu_kl_F"]) < 5e-2 and relerr(Wl, gold["mu_kl_W"]) < 5e-2


def test_kmeans_and_projection(gold):
    from gpzoo_b200 import initialisation as I, kernels
    g = torch.Generator().manual_seed(4)
    centres = torch.tensor([[-3.0, 0.0], [3.0, 1.0], [0.0, 4.0], [1.0, -4.0], [5.0, 5.0]])
    X = (centres[torch.randint(5, (2000,), generator=g)] + 0.3 * torch.randn(2000, 2, generator=g)).to(DEV)
    Z, inertia = I.kmeans_inducing(X, 5, seed=1)
    d = torch.cdist(Z.cpu(), centres).min(1).values          # every centre is recovered
    assert float(d.max()) < 0.1 and abs(inertia / (2000 * 2 * 0.09) - 1) < 0.1
    Z64, inertia64 = I.kmeans_inducing(X.double().cpu().numpy(), 40, seed=2)
    assert Z64.shape == (40, 2) and isinstance(Z64, np.ndarray) and inertia64 < inertia
    # projection of log-scale factors onto inducing points (Slideseqv2_estimate_lengthscales.ipynb cell 16), fp64
    kern = kernels.NSF_RBF(L=3, sigma=1.0, lengthscale=1.0).to(DEV).double()
    mu = I.project_to_inducing(kern, torch.from_numpy(gold["proj_Z"]).to(DEV), torch.from_numpy(gold["coords_out"]).to(DEV),
                               torch.from_numpy(gold["proj_factors"]).to(DEV), jitter=1e-5)
    assert relerr(mu, gold["proj_mu"]) < 1e-6
