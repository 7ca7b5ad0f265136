"""CPU oracle for the GPzoo sparse-GP ELBO hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the arithmetic that the reference
executes on the hot path, written as plain functions over CPU torch tensors so
that (a) autograd gives reference gradients, (b) the op sequence is the one the
reference issues (cdist -> exp -> cholesky -> cholesky_solve -> W@(S-Kzz) ...),
which makes it a fair CPU timing baseline (`bench.py` cpu_baseline, kind="port").

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module.  The product package
`gpzoo_b200` never does.

Parity pinning: the reference repository contains no golden vectors or tests
(SURVEY.md §4, §8c), so this oracle is pinned against *outputs of the reference
itself*, generated in the build container by `oracle/gen_golden.py` (which
imports the unmodified reference from /root/reference) and committed under
`tests/golden/`.  `tests/test_oracle.py` checks oracle == golden to ~1e-12
(fp64) and, when /root/reference is present, oracle == live reference.

Arithmetic lives in PyTorch (un-pinned by the reference's setup.py:4-7); the
oracle pins it to torch 2.11.0 as installed in this image.

All citations `file:line` are into /root/reference/gpzoo/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn.functional as Fnn

# --------------------------------------------------------------------------
# L0: utilities.py hot-path helpers
# --------------------------------------------------------------------------


def squared_dist(X, Z):
    """|x|^2 - 2 x.z + |z|^2 clamped at 0   (utilities.py:399-405)."""
    x2 = (X * X).sum(1, keepdim=True)
    z2 = (Z * Z).sum(1, keepdim=True)
    return (x2 - 2.0 * (X @ Z.t()) + z2.t()).clamp(min=0)


def embed_distance_matrix(D):
    """Classical MDS embedding of a group-distance matrix (utilities.py:459-469)."""
    n = D.shape[0]
    C = torch.eye(n, dtype=D.dtype, device=D.device) - torch.ones(n, n, dtype=D.dtype, device=D.device) / n
    B = -0.5 * (C @ (D * D) @ C)
    lam, Q = torch.linalg.eigh(B)
    lam = torch.where(lam < 0, torch.zeros_like(lam), lam)
    return Q @ torch.diag((lam + 1e-6).sqrt())       # _torch_sqrt(., 1e-6)  utilities.py:450-456


def with_jitter(K, jitter):
    """add_jitter (utilities.py:407-418) adds `jitter` to the diagonal IN PLACE and the
    jittered matrix is what flows on; the out-of-place form has the same value/gradient."""
    M = K.shape[-1]
    return K + jitter * torch.eye(M, dtype=K.dtype, device=K.device)


def lower_cholesky_transform(raw):
    """transform_to(constraints.lower_cholesky) (gp.py:50,220,369; torch transforms.py _call)."""
    return raw.tril(-1) + torch.diag_embed(torch.diagonal(raw, dim1=-2, dim2=-1).exp())


def svgp_forward(Kxx, Kzz, W, mu, S):
    """mean = W mu ; cov = Kxx + sum_j (W (S-Kzz)) o W      (utilities.py:382-397)."""
    mean = W @ mu.unsqueeze(-1)
    cov = Kxx + ((W @ (S - Kzz)) * W).sum(-1)
    return mean, cov


def whitened_kl(mz, Lz):
    """utilities.py:27-36 (single GP: Lz is M x M)."""
    M = mz.shape[-1]
    return 0.5 * (-2.0 * torch.diagonal(Lz).log().sum() + (Lz * Lz).sum() + (mz * mz).sum() - M)


# --------------------------------------------------------------------------
# L1: kernels.py
# --------------------------------------------------------------------------

_CDIST_MODE = "use_mm_for_euclid_dist_if_necessary"   # torch.cdist default (kernels.py:118,146)


def set_exact_cdist(flag: bool):
    """torch.cdist's matmul path costs ~1e-4 relative accuracy at +-100 coordinates.
    `flag=True` forces the direct-difference path (same maths, tighter fp32)."""
    global _CDIST_MODE
    _CDIST_MODE = "donot_use_mm_for_euclid_dist" if flag else "use_mm_for_euclid_dist_if_necessary"


def _cdist(X, Z):
    return torch.cdist(X, Z, compute_mode=_CDIST_MODE)


def rbf(X, Z, sigma, lengthscale):
    """RBF.forward (kernels.py:114-130): scalar sigma, lengthscale -> |X| x |Z|."""
    d = _cdist(X, Z)
    return sigma ** 2 * torch.exp(-0.5 * d ** 2 / lengthscale ** 2)


def nsf_rbf(X, Z, sigma, lengthscale):
    """NSF_RBF.forward (kernels.py:141-155): sigma, lengthscale of shape (L,1,1) -> L x |X| x |Z|."""
    d2 = (_cdist(X, Z) ** 2)[None]
    return sigma ** 2 * torch.exp(-0.5 * d2 / lengthscale ** 2)


def rbf_diag(X, sigma):
    """diag=True branches (kernels.py:115-116, 143-144)."""
    if sigma.dim() == 0:
        return (sigma ** 2).expand(X.shape[0])
    return (sigma ** 2).reshape(-1)[:, None].expand(-1, X.shape[0])


def mggp_rbf(X, Z, gX, gZ, sigma, lengthscale, gdp, embedding, input_dim=2):
    """MGGP_RBF.forward (kernels.py:172-191): a = gdp (not squared)."""
    r2 = squared_dist(embedding[gX], embedding[gZ])
    d2 = squared_dist(X, Z) / lengthscale ** 2
    den = gdp * r2 + 1
    return sigma ** 2 * torch.exp(-0.5 * d2 / den) / den ** (0.5 * input_dim)


def mggp_nsf_rbf(X, Z, gX, gZ, sigma, lengthscale, gdp, embedding, input_dim=2):
    """MGGP_NSF_RBF.forward (kernels.py:204-228): a = gdp**2, params (L,1,1)."""
    r2 = squared_dist(embedding[gX], embedding[gZ])[None]
    d2 = squared_dist(X, Z)[None] / lengthscale ** 2
    den = gdp ** 2 * r2 + 1
    return sigma ** 2 * torch.exp(-0.5 * d2 / den) / den ** (0.5 * input_dim)


def matern32(X, Z, sigma, lengthscale):
    """batched_Matern32.covariance (kernels.py:14-20), scalar params -> |X| x |Z|."""
    d = ((X[:, None, :] - Z[None, :, :]) ** 2).sum(-1).sqrt()
    v = math.sqrt(3.0) * d / lengthscale
    return sigma ** 2 * (1 + v) * torch.exp(-v)


# --------------------------------------------------------------------------
# L2: gp.py
# --------------------------------------------------------------------------


def svgp(Kxx, Kzx, Kzz, mu, Lu_raw, jitter, clamp_min):
    """SVGP.forward / MGGP_SVGP.forward after the kernel calls (gp.py:208-230, 360-380).

    Returns mean (L x N), clamped variance (L x N), Lu, Lc."""
    Kzz = with_jitter(Kzz.contiguous(), jitter)                   # gp.py:208-209 (in-place add_jitter)
    Lc = torch.linalg.cholesky(Kzz)                               # gp.py:213
    W = torch.cholesky_solve(Kzx, Lc).transpose(-2, -1)           # gp.py:218-219
    Lu = lower_cholesky_transform(Lu_raw)                         # gp.py:220
    S = Lu @ Lu.transpose(-2, -1)                                 # gp.py:221
    mean, cov = svgp_forward(Kxx, Kzz, W, mu, S)                  # gp.py:225
    mean = mean.squeeze(-1)
    var = torch.clamp(cov, min=clamp_min)                         # gp.py:228 (1e-6) / :378 (5e-2)
    return mean, var, Lu, Lc


def wsvgp(Kxx, Kzx, Kzz, mu, Lu_raw, jitter):
    """WSVGP.forward (gp.py:260-306)."""
    Kzz = with_jitter(Kzz.contiguous(), jitter)
    Lc = torch.linalg.cholesky(Kzz)
    W = torch.linalg.solve_triangular(Lc, Kzx, upper=False).transpose(-2, -1)   # gp.py:276-277
    Lu = lower_cholesky_transform(Lu_raw)
    cov = torch.clamp(Kxx - (W ** 2).sum(-1), min=0.0) + ((W @ Lu) ** 2).sum(-1)  # gp.py:286-288
    mean = (W @ mu.unsqueeze(-1)).squeeze(-1)
    return mean, cov, Lu


def vnngp(Kxx, Kxz, dist_xz, Kzz, mu, Lu_raw, jitter, K, clamp_min=5e-2):
    """VNNGP.forward (gp.py:19-122).  Kxx: L x N, Kxz: L x N x M (note X,Z orientation gp.py:31),
    dist_xz: N x M (un-batched cdist), Kzz: L x M x M.  Returns mean, var, Lu, Lc, indexes."""
    L, N, M = Kxz.shape
    Lu = lower_cholesky_transform(Lu_raw).reshape(-1, M, M)
    Lc = torch.linalg.cholesky(with_jitter(Kzz, jitter))                    # gp.py:55
    nn = torch.argsort(dist_xz, dim=1)[:, :K]                               # gp.py:64
    lL = Lc[:, nn]                                                          # L x N x K x M  gp.py:67
    kzz = (lL @ lL.transpose(-2, -1)).reshape(-1, K, K)                     # gp.py:72-74
    kzz = with_jitter(kzz, jitter)                                          # second jitter gp.py:77
    kinv = torch.inverse(kzz)
    kxz = torch.gather(Kxz.reshape(-1, M), 1, nn.repeat(L, 1))[:, None, :]  # gp.py:83-86
    W = kxz @ kinv                                                          # gp.py:88
    lmu = mu.reshape(-1, M)[:, nn].reshape(-1, K)                           # gp.py:97-98
    lLu = Lu[:, nn]
    lS = (lLu @ lLu.transpose(-2, -1)).reshape(-1, K, K)                    # gp.py:100-102
    mean, cov = svgp_forward(Kxx.reshape(-1, 1), kzz, W, lmu, lS)           # gp.py:106
    mean = mean.reshape(L, N)
    cov = cov.reshape(L, N)
    return mean, torch.clamp(cov, min=clamp_min), Lu, Lc, nn


def mvn_kl(mu, Lu, Lc):
    """kl_divergence(MVN(mu, Lu), MVN(0, Lc)) as torch computes it (torch/distributions/kl.py
    _kl_multivariatenormal_multivariatenormal; called at utilities.py:481,616).  -> (L,)"""
    M = mu.shape[-1]
    half = torch.diagonal(Lc, dim1=-2, dim2=-1).log().sum(-1) - torch.diagonal(Lu, dim1=-2, dim2=-1).log().sum(-1)
    P = torch.linalg.solve_triangular(Lc, Lu, upper=False)
    q = torch.linalg.solve_triangular(Lc, mu.unsqueeze(-1), upper=False)
    return half + 0.5 * ((P * P).sum((-2, -1)) + (q * q).sum((-2, -1)) - M)


def normal_kl(m, s, m0, s0):
    """Normal||Normal KL (torch kl.py _kl_normal_normal; utilities.py:515-516)."""
    vr = (s / s0) ** 2
    return 0.5 * (vr + ((m - m0) / s0) ** 2 - 1 - vr.log())


# --------------------------------------------------------------------------
# L3: likelihoods.py
# --------------------------------------------------------------------------


def poisson_rate(W, F, V, softplus_W=True):
    """rate[e,g,n] = softplus(V_n) * sum_l softplus(W_gl) exp(F_eln)
    (likelihoods.py:49-53, 83-85; Hybrid_NSF uses raw W: likelihoods.py:293,322)."""
    Wp = Fnn.softplus(W) if softplus_W else W
    return Fnn.softplus(V) * torch.matmul(Wp, torch.exp(F))


def poisson_loglik(y, rate, with_lgamma=True):
    """Poisson.log_prob (torch poisson.py:75-79) or the notebooks' y*log(rate)-rate
    (utilities.py:507; Slideseq_NSF_newest_version.ipynb:411).  Returns mean_E sum_{g,n}."""
    lp = torch.xlogy(y, rate) - rate
    if with_lgamma:
        lp = lp - torch.lgamma(y + 1)
    return lp.mean(0).sum()


def gaussian_loglik(y, F, noise_raw):
    """GaussianLikelihood (likelihoods.py:14-20) + train() ELBO term (utilities.py:479)."""
    s = Fnn.softplus(noise_raw)
    lp = -((y - F) ** 2) / (2 * s ** 2) - s.log() - 0.5 * math.log(2 * math.pi)
    return lp.mean(0).sum()


# --------------------------------------------------------------------------
# End-to-end ELBOs (the "step" that is timed): utilities.py:471-493, 600-631
# --------------------------------------------------------------------------


@dataclass
class NSFParams:
    Z: torch.Tensor          # M x D
    sigma: torch.Tensor      # L x 1 x 1
    lengthscale: torch.Tensor  # L x 1 x 1
    mu: torch.Tensor         # L x M
    Lu_raw: torch.Tensor     # L x M x M
    W: torch.Tensor          # G x L
    V: torch.Tensor          # N
    jitter: float = 1e-1
    # multi-group extras (None for plain NSF_RBF)
    gdp: Optional[torch.Tensor] = None
    embedding: Optional[torch.Tensor] = None
    groupsZ: Optional[torch.Tensor] = None

    def leaves(self):
        out = dict(Z=self.Z, sigma=self.sigma, lengthscale=self.lengthscale, mu=self.mu,
                   Lu_raw=self.Lu_raw, W=self.W, V=self.V)
        if self.gdp is not None:
            out["gdp"] = self.gdp
        return out


def nsf_svgp_terms(p: NSFParams, X, y, eps, idx=None, groupsX=None, with_lgamma=True, clamp_min=None):
    """NSF2(SVGP(NSF_RBF)) / NSF2(MGGP_SVGP(MGGP_NSF_RBF)) forward[_batched] + ELBO pieces.

    Call stack restated: likelihoods.py:80-97 -> gp.py:183-232 (or 341-382) -> kernels.py:141-155
    (or 204-228) -> utilities.py:382-418 -> utilities.py:611-616.
    eps: E x L x B standard normal (the draw inside Normal.rsample, torch normal.py rsample)."""
    Xb = X if idx is None else X[idx]
    yb = y if idx is None else y[:, idx]
    Vb = p.V if idx is None else p.V[idx]
    if p.gdp is None:
        Kzx = nsf_rbf(p.Z, Xb, p.sigma, p.lengthscale)
        Kzz = nsf_rbf(p.Z, p.Z, p.sigma, p.lengthscale)
        cmin = 1e-6 if clamp_min is None else clamp_min
    else:
        gXb = groupsX if idx is None else groupsX[idx]
        Kzx = mggp_nsf_rbf(p.Z, Xb, p.groupsZ, gXb, p.sigma, p.lengthscale, p.gdp, p.embedding)
        Kzz = mggp_nsf_rbf(p.Z, p.Z, p.groupsZ, p.groupsZ, p.sigma, p.lengthscale, p.gdp, p.embedding)
        cmin = 5e-2 if clamp_min is None else clamp_min
    Kxx = rbf_diag(Xb, p.sigma)
    mean, var, Lu, Lc = svgp(Kxx, Kzx, Kzz, p.mu, p.Lu_raw, p.jitter, cmin)
    sd = var.sqrt()
    F = mean + eps * sd                                             # qF.rsample((E,))
    rate = poisson_rate(p.W, F, Vb)
    ll = poisson_loglik(yb, rate, with_lgamma)
    kl = mvn_kl(p.mu, Lu, Lc)
    return dict(elbo=ll - kl.sum(), ll=ll, kl=kl, mean=mean, var=var, Lu=Lu, Lc=Lc, rate=rate)


def svgp_gaussian_terms(Z, sigma, lengthscale, mu, Lu_raw, noise_raw, X, y, eps, jitter):
    """GaussianLikelihood(SVGP(RBF)) (config 1): likelihoods.py:14-20, gp.py:183-232, kernels.py:114-130."""
    Kzx = rbf(Z, X, sigma, lengthscale)
    Kzz = rbf(Z, Z, sigma, lengthscale)
    Kxx = rbf_diag(X, sigma)
    mean, var, Lu, Lc = svgp(Kxx, Kzx, Kzz, mu, Lu_raw, jitter, 1e-6)
    F = mean + eps * var.sqrt()
    ll = gaussian_loglik(y, F, noise_raw)
    kl = mvn_kl(mu, Lu, Lc)
    return dict(elbo=ll - kl.sum(), ll=ll, kl=kl, mean=mean, var=var, Lu=Lu, Lc=Lc)


def hybrid_terms(p: NSFParams, Wcf, cf_mean, cf_scale_raw, X, y, eps1, eps2, idx=None,
                 scale_pf=1.0, with_lgamma=True):
    """Hybrid_NSF2(SVGP, GaussianPrior) forward[_batched] (likelihoods.py:110-145; gp.py:125-146)
    and the two-KL ELBO (utilities.py:509-516)."""
    Xb = X if idx is None else X[idx]
    yb = y if idx is None else y[:, idx]
    Vb = p.V if idx is None else p.V[idx]
    Kzx = nsf_rbf(p.Z, Xb, p.sigma, p.lengthscale)
    Kzz = nsf_rbf(p.Z, p.Z, p.sigma, p.lengthscale)
    Kxx = rbf_diag(Xb, p.sigma)
    mean, var, Lu, Lc = svgp(Kxx, Kzx, Kzz, p.mu, p.Lu_raw, p.jitter, 1e-6)
    m2 = cf_mean if idx is None else cf_mean[:, idx]
    s2 = Fnn.softplus(cf_scale_raw if idx is None else cf_scale_raw[:, idx])
    F1 = mean + eps1 * var.sqrt()
    F2 = m2 + eps2 * s2
    Zr = torch.matmul(Fnn.softplus(p.W), torch.exp(F1)) + torch.matmul(Fnn.softplus(Wcf), torch.exp(F2))
    rate = Fnn.softplus(Vb) * Zr
    ll = poisson_loglik(yb, rate, with_lgamma)
    kl = mvn_kl(p.mu, Lu, Lc)
    kl2 = normal_kl(m2, s2, torch.zeros_like(m2), scale_pf * torch.ones_like(s2))
    return dict(elbo=ll - kl.sum() - kl2.sum(), ll=ll, kl=kl, kl2=kl2, mean=mean, var=var, rate=rate)


def vnngp_terms(p: NSFParams, X, y, eps, K, with_lgamma=True):
    """NSF2(VNNGP(NSF_RBF)) (config 3): likelihoods.py:80-87 -> gp.py:19-122."""
    Kxx = rbf_diag(X, p.sigma)
    dist = _cdist(X, p.Z)
    Kxz = p.sigma ** 2 * torch.exp(-0.5 * (dist ** 2)[None] / p.lengthscale ** 2)
    Kzz = nsf_rbf(p.Z, p.Z, p.sigma, p.lengthscale)
    mean, var, Lu, Lc, nn = vnngp(Kxx, Kxz, dist, Kzz, p.mu, p.Lu_raw, p.jitter, K)
    F = mean + eps * var.sqrt()
    rate = poisson_rate(p.W, F, p.V)
    ll = poisson_loglik(y, rate, with_lgamma)
    kl = mvn_kl(p.mu, Lu, Lc)
    return dict(elbo=ll - kl.sum(), ll=ll, kl=kl, mean=mean, var=var, nn=nn)


def value_and_grads(fn: Callable[[], dict], leaves: dict):
    """Run `fn` (which must use the tensors in `leaves`), back-propagate -ELBO as the reference's
    training loops do (utilities.py:484-485, 619-620) and return (terms, {name: d ELBO/d leaf})."""
    for t in leaves.values():
        t.requires_grad_(True)
        t.grad = None
    out = fn()
    out["elbo"].backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    for t in leaves.values():
        t.grad = None
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, grads


# ------------------------------------------------------------------------------------------------
# initialisation pipeline: the deterministic part of utilities.py:237-313, restated on numpy-like arrays
# ------------------------------------------------------------------------------------------------
def lnormal_approx_dirichlet(L):
    """utilities.py:237-250."""
    import math
    s2 = math.log(2 * L) - math.log(L + 1)
    return -math.log(L) - s2 / 2.0, math.sqrt(s2)


def regularized_nmf_post(eF, W, L, sz=1.0, pseudocount=1e-2, shrinkage=0.2):
    """What regularized_nmf does to the NMF's (factors N x L, loadings G x L): utilities.py:284-299 with the shrinkage of
    :301-313.  torch tensors (fp64) in and out."""
    a = shrinkage
    if 0 < a < 1:
        W = W * (1 - a) + a * W.sum(0) / float(W.shape[0])
    wsum = W.sum(0)
    eF = eF * wsum
    if 0 < a < 1:
        eF = eF * (1 - a) + a * eF.sum(1, keepdim=True) / float(eF.shape[1])
    Fl = torch.log(pseudocount + eF) - torch.log(torch.as_tensor(sz, dtype=eF.dtype))
    mu, _ = lnormal_approx_dirichlet(max(L, 1.1))
    wt = Fl.mean(0) - mu
    return Fl - wt, W * torch.exp(wt - torch.log(wsum))


def init_softplus(mat, minval=1e-5):
    """utilities.py:38-44."""
    return torch.where(mat < 20, torch.log(torch.exp(mat.clamp(max=20.0)) - 1 + minval), mat)


def rescale_spatial_coords(X, box_side=4):
    """utilities.py:71-84."""
    X = X - X.min(0).values
    X = X * (box_side / torch.exp(torch.log(X.max(0).values).mean()))
    return X - X.mean(0)
