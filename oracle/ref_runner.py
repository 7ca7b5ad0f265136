"""Drive the UNMODIFIED reference modules on a synthetic problem and collect ELBO pieces and
gradients.  TEST INFRASTRUCTURE ONLY; requires /root/reference (build container only).

Used by oracle/gen_golden.py (to write tests/golden/*.npz) and by tests/test_oracle.py (live
oracle-vs-reference comparison when the reference is present).
"""
import torch
from torch import distributions, nn

from .ref_loader import fixed_eps, load_reference, quiet


def _p(t):
    return nn.Parameter(t.clone())


def _grads(named):
    return {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in named.items()}


def _elbo(pY, y, qU, pU, with_lgamma=True):
    # utilities.py:479-481 / 611-616 (log_prob form) or :507 (y*log(rate)-rate form)
    if with_lgamma:
        ll = pY.log_prob(y).mean(axis=0).sum()
    else:
        ll = (y * torch.log(pY.rate) - pY.rate).mean(axis=0).sum()
    kl = distributions.kl_divergence(qU, pU)
    return ll, kl


def run_nsf_svgp(prob, idx=None, with_lgamma=True):
    """NSF2(SVGP(NSF_RBF)) or, if prob has groups, NSF2(MGGP_SVGP(MGGP_NSF_RBF))."""
    ref = load_reference()
    L, M = prob["mu"].shape
    D = prob["X"].shape[1]
    dt = prob["X"].dtype
    mg = "groupsX" in prob
    if mg:
        ng = prob["group_distances"].shape[0]
        kern = ref.kernels.MGGP_NSF_RBF(L=L, n_groups=ng)
        kern.embedding = nn.Parameter(ref.utilities._embed_distance_matrix(prob["group_distances"].float()).to(dt),
                                      requires_grad=False)
        kern.group_diff_param = _p(prob["gdp"])
        gp = ref.gp.MGGP_SVGP(kern, dim=D, M=M, jitter=prob["jitter"], n_groups=ng)
        gp.groupsZ = nn.Parameter(prob["groupsZ"].clone(), requires_grad=False)
    else:
        kern = ref.kernels.NSF_RBF(L=L)
        gp = ref.gp.SVGP(kern, dim=D, M=M, jitter=prob["jitter"])
    kern.sigma = _p(prob["sigma"])
    kern.lengthscale = _p(prob["lengthscale"])
    gp.Z = _p(prob["Z"])
    gp.mu = _p(prob["mu"])
    gp.Lu = _p(prob["Lu_raw"])
    model = ref.likelihoods.NSF2(gp, prob["y"], L=L)
    model.W = _p(prob["W"])
    model.V = _p(prob["V"])
    kw = {}
    with fixed_eps(prob["eps"] if idx is None else prob["eps"][:, :, idx]):
        if idx is None:
            if mg:
                kw["groupsX"] = prob["groupsX"]
            pY, qF, qU, pU = model(X=prob["X"], E=prob["eps"].shape[0], **kw)
            yb = prob["y"]
        else:
            if mg:
                kw["groupsX"] = prob["groupsX"][idx]
            pY, qF, qU, pU = model.forward_batched(X=prob["X"], idx=idx, E=prob["eps"].shape[0], **kw)
            yb = prob["y"][:, idx]
    ll, kl = _elbo(pY, yb, qU, pU, with_lgamma)
    elbo = ll - kl.sum()
    (-elbo).backward()
    named = dict(Z=gp.Z, sigma=kern.sigma, lengthscale=kern.lengthscale, mu=gp.mu, Lu_raw=gp.Lu,
                 W=model.W, V=model.V)
    if mg:
        named["gdp"] = kern.group_diff_param
    grads = {k: -v for k, v in _grads(named).items()}          # d ELBO / d param
    out = dict(elbo=elbo, ll=ll, kl=kl, mean=qF.mean, var=qF.scale ** 2, Lu=qU.scale_tril, Lc=pU.scale_tril)
    if mg:
        out["embedding"] = kern.embedding
    return {k: v.detach() for k, v in out.items()}, grads


def run_svgp_gaussian(prob):
    """GaussianLikelihood(SVGP(RBF)) — config 1."""
    ref = load_reference()
    M = prob["mu"].shape[0]
    kern = ref.kernels.RBF()
    kern.sigma = _p(prob["sigma"])
    kern.lengthscale = _p(prob["lengthscale"])
    gp = ref.gp.SVGP(kern, dim=prob["X"].shape[1], M=M, jitter=prob["jitter"])
    gp.Z, gp.mu, gp.Lu = _p(prob["Z"]), _p(prob["mu"]), _p(prob["Lu_raw"])
    model = ref.likelihoods.GaussianLikelihood(gp)
    model.noise = _p(prob["noise"])
    with fixed_eps(prob["eps"]):
        pY, qF, qU, pU = model(X=prob["X"], E=prob["eps"].shape[0])
    ll, kl = _elbo(pY, prob["y"], qU, pU)
    elbo = ll - kl.sum()
    (-elbo).backward()
    named = dict(Z=gp.Z, sigma=kern.sigma, lengthscale=kern.lengthscale, mu=gp.mu, Lu_raw=gp.Lu, noise=model.noise)
    grads = {k: -v for k, v in _grads(named).items()}
    out = dict(elbo=elbo, ll=ll, kl=kl, mean=qF.mean, var=qF.scale ** 2)
    return {k: v.detach() for k, v in out.items()}, grads


def run_vnngp(prob, K):
    """NSF2(VNNGP(NSF_RBF)) — config 3."""
    ref = load_reference()
    L, M = prob["mu"].shape
    kern = ref.kernels.NSF_RBF(L=L)
    kern.sigma, kern.lengthscale = _p(prob["sigma"]), _p(prob["lengthscale"])
    gp = ref.gp.VNNGP(kern, dim=prob["X"].shape[1], M=M, K=K, jitter=prob["jitter"])
    gp.Z, gp.mu, gp.Lu = _p(prob["Z"]), _p(prob["mu"]), _p(prob["Lu_raw"])
    model = ref.likelihoods.NSF2(gp, prob["y"], L=L)
    model.W, model.V = _p(prob["W"]), _p(prob["V"])
    with fixed_eps(prob["eps"]), quiet():
        pY, qF, qU, pU = model(X=prob["X"], E=prob["eps"].shape[0])
        nn_idx = torch.argsort(torch.cdist(prob["X"], gp.Z.detach()), dim=1)[:, :K]   # gp.py:64
    ll, kl = _elbo(pY, prob["y"], qU, pU)
    elbo = ll - kl.sum()
    (-elbo).backward()
    named = dict(Z=gp.Z, sigma=kern.sigma, lengthscale=kern.lengthscale, mu=gp.mu, Lu_raw=gp.Lu,
                 W=model.W, V=model.V)
    grads = {k: -v for k, v in _grads(named).items()}
    out = dict(elbo=elbo, ll=ll, kl=kl, mean=qF.mean, var=qF.scale ** 2, nn=nn_idx)
    return {k: v.detach() for k, v in out.items()}, grads


def run_hybrid(prob, extra, idx):
    """Hybrid_NSF2(SVGP(NSF_RBF), GaussianPrior) forward_batched — config 5 shape."""
    ref = load_reference()
    L, M = prob["mu"].shape
    T = extra["Wcf"].shape[1]
    kern = ref.kernels.NSF_RBF(L=L)
    kern.sigma, kern.lengthscale = _p(prob["sigma"]), _p(prob["lengthscale"])
    gp = ref.gp.SVGP(kern, dim=prob["X"].shape[1], M=M, jitter=prob["jitter"])
    gp.Z, gp.mu, gp.Lu = _p(prob["Z"]), _p(prob["mu"]), _p(prob["Lu_raw"])
    prior = ref.gp.GaussianPrior(prob["y"], L=T)
    prior.mean, prior.scale = _p(extra["cf_mean"]), _p(extra["cf_scale"])
    model = ref.likelihoods.Hybrid_NSF2(gp, prior, prob["y"], L=L, T=T)
    model.sf.W, model.cf.W, model.V = _p(prob["W"]), _p(extra["Wcf"]), _p(prob["V"])
    with fixed_eps(prob["eps"][:, :, idx], extra["eps2"][:, :, idx]):
        pY, qF, qU, pU, qF2, pF2 = model.forward_batched(X=prob["X"], idx=idx, E=prob["eps"].shape[0])
    ll, kl = _elbo(pY, prob["y"][:, idx], qU, pU)
    kl2 = distributions.kl_divergence(qF2, pF2)
    elbo = ll - kl.sum() - kl2.sum()
    (-elbo).backward()
    named = dict(Z=gp.Z, sigma=kern.sigma, lengthscale=kern.lengthscale, mu=gp.mu, Lu_raw=gp.Lu,
                 W=model.sf.W, V=model.V, Wcf=model.cf.W, cf_mean=prior.mean, cf_scale=prior.scale)
    grads = {k: -v for k, v in _grads(named).items()}
    out = dict(elbo=elbo, ll=ll, kl=kl, kl2=kl2.sum(), mean=qF.mean, var=qF.scale ** 2)
    return {k: v.detach() for k, v in out.items()}, grads


def run_wsvgp(prob):
    """NSF2(WSVGP(NSF_RBF)) or, with groups, NSF2(MGGP_WSVGP(MGGP_NSF_RBF)): forward (gp.py:260-306, 385-399), the ELBO a user
    assembles with whitened_KL per factor (utilities.py:27-36; pZ is None so kl_divergence is never called), its gradients, and
    forward_precomputed (gp.py:308-322) on W = Kxz Lc^-T with a batched_RBF-style (L,) sigma."""
    ref = load_reference()
    L, M = prob["mu"].shape
    D = prob["X"].shape[1]
    dt = prob["X"].dtype
    mg = "groupsX" in prob
    if mg:
        ng = prob["group_distances"].shape[0]
        kern = ref.kernels.MGGP_NSF_RBF(L=L, n_groups=ng)
        kern.embedding = nn.Parameter(ref.utilities._embed_distance_matrix(prob["group_distances"].float()).to(dt),
                                      requires_grad=False)
        kern.group_diff_param = _p(prob["gdp"])
        gp = ref.gp.MGGP_WSVGP(kern, dim=D, M=M, n_groups=ng, jitter=prob["jitter"])
        gp.groupsZ = nn.Parameter(prob["groupsZ"].clone(), requires_grad=False)
        kw = dict(groupsX=prob["groupsX"])
    else:
        kern = ref.kernels.NSF_RBF(L=L)
        gp = ref.gp.WSVGP(kern, dim=D, M=M, jitter=prob["jitter"])
        kw = {}
    kern.sigma, kern.lengthscale = _p(prob["sigma"]), _p(prob["lengthscale"])
    gp.Z, gp.mu, gp.Lu = _p(prob["Z"]), _p(prob["mu"]), _p(prob["Lu_raw"])
    model = ref.likelihoods.NSF2(gp, prob["y"], L=L)
    model.W, model.V = _p(prob["W"]), _p(prob["V"])
    with fixed_eps(prob["eps"]):
        pY, qF, qZ, pZ = model(X=prob["X"], E=prob["eps"].shape[0], **kw)
    assert pZ is None
    ll = pY.log_prob(prob["y"]).mean(axis=0).sum()
    kl = torch.stack([ref.utilities.whitened_KL(gp.mu[l], qZ.scale_tril[l]) for l in range(L)])
    elbo = ll - kl.sum()
    (-elbo).backward()
    named = dict(Z=gp.Z, sigma=kern.sigma, lengthscale=kern.lengthscale, mu=gp.mu, Lu_raw=gp.Lu, W=model.W, V=model.V)
    if mg:
        named["gdp"] = kern.group_diff_param
    grads = {k: -v for k, v in _grads(named).items()}
    out = dict(elbo=elbo, ll=ll, kl=kl, mean=qF.mean, var=qF.scale ** 2, Lu=qZ.scale_tril)
    with torch.no_grad():
        Kxx, Kzx, Kzz = gp.forward_kernels(prob["X"], **kw)
        Lc = torch.linalg.cholesky(ref.utilities.add_jitter(Kzz.clone(), prob["jitter"]))
        Wp = torch.linalg.solve_triangular(Lc, Kzx, upper=False).transpose(-2, -1).contiguous()          # L x N x M
        if not mg:
            kb = ref.kernels.batched_RBF()
            kb.sigma, kb.lengthscale = _p(prob["sigma"].reshape(-1)), _p(prob["lengthscale"].reshape(-1))
            gpb = ref.gp.WSVGP(kb, dim=D, M=M, jitter=prob["jitter"])
            gpb.Z, gpb.mu, gpb.Lu = gp.Z, gp.mu, gp.Lu
            qFp, _, _ = gpb.forward_precomputed(Wp)
            out["pre_mean"], out["pre_var"] = qFp.mean, qFp.scale ** 2
        out["pre_W"] = Wp
    return {k: v.detach() for k, v in out.items()}, grads


def model_zoo(ns, N=12, G=5, M=6, L=3, T=2, ng=3):
    """One small instance of every model family, built from `ns.kernels / ns.gp / ns.likelihoods` (the reference or gpzoo_b200)
    with the constructor calls the notebooks use.  Used to compare state_dict keys and shapes."""
    y = torch.zeros(G, N)
    k, g, lk = ns.kernels, ns.gp, ns.likelihoods
    zoo = {
        "GaussianLikelihood(SVGP(RBF))": lk.GaussianLikelihood(g.SVGP(k.RBF(), dim=2, M=M, jitter=1e-3)),
        "NSF2(SVGP(NSF_RBF))": lk.NSF2(g.SVGP(k.NSF_RBF(L=L), dim=2, M=M, jitter=1e-1), y, L=L),
        "NSF2(VNNGP(NSF_RBF))": lk.NSF2(g.VNNGP(k.NSF_RBF(L=L), dim=2, M=M, K=3, jitter=1e-2), y, L=L),
        "NSF2(MGGP_SVGP(MGGP_NSF_RBF))": lk.NSF2(g.MGGP_SVGP(k.MGGP_NSF_RBF(L=L, n_groups=ng), dim=2, M=M, jitter=1e-1, n_groups=ng), y, L=L),
        "MGGP_NSF(MGGP_SVGP(MGGP_RBF))": lk.MGGP_NSF(g.MGGP_SVGP(k.MGGP_RBF(n_groups=ng), dim=2, M=M, jitter=1e-1, n_groups=ng), y, L=L),
        "NSF(SVGP(RBF))": lk.NSF(g.SVGP(k.RBF(), dim=2, M=M), y, L=L),
        "Hybrid_NSF(SVGP(RBF))": lk.Hybrid_NSF(g.SVGP(k.RBF(), dim=2, M=M), y, L=L),
        "Hybrid_NSF2(SVGP(NSF_RBF),GaussianPrior)": lk.Hybrid_NSF2(g.SVGP(k.NSF_RBF(L=L), dim=2, M=M), g.GaussianPrior(y, L=T), y, L=L, T=T),
        "PNMF(GaussianPrior)": lk.PNMF(g.GaussianPrior(y, L=L), y, L=L),
        "NSF2(WSVGP(NSF_RBF))": lk.NSF2(g.WSVGP(k.NSF_RBF(L=L), dim=2, M=M), y, L=L),
        "NSF2(MGGP_WSVGP(batched_MGGP_RBF))": lk.NSF2(g.MGGP_WSVGP(k.batched_MGGP_RBF(n_groups=ng), dim=2, M=M, n_groups=ng), y, L=L),
        "SVGP(batched_RBF)": g.SVGP(k.batched_RBF(), dim=2, M=M),
        "SVGP(batched_Matern32)": g.SVGP(k.batched_Matern32(), dim=2, M=M),
    }
    return zoo


def state_dict_shapes(ns=None):
    ns = load_reference() if ns is None else ns
    return {name: {key: list(v.shape) for key, v in m.state_dict().items()} for name, m in model_zoo(ns).items()}
