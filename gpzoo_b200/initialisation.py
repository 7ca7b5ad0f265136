"""Initialisation pipeline on the GPU: what the reference's notebooks run on the host before training.

    regularized_nmf / shrink_factors / shrink_loadings / lnormal_approx_dirichlet   utilities.py:237-313 (NSF-paper code)
    init_softplus, rescale_spatial_coords, build_group_distances                    utilities.py:14-24, 38-44, 71-84
    kmeans_inducing                 sklearn KMeans(...).cluster_centers_            Slideseqv2_estimate_lengthscales.ipynb cell 16
    project_to_inducing             mu = Kzz (Kzx Kxz + j I)^-1 Kzx F               idem (and NSF_Hybrid_benchmark.ipynb cell 11)

The reference hands the factorisation to `sklearn.decomposition.NMF` on the host (minutes at Slide-seq size: every iteration makes
several passes over the N x G count matrix).  Here the same two solvers run on the device: multiplicative updates for the
Frobenius / Kullback-Leibler losses (sklearn `_fit_multiplicative_update`) and cyclic coordinate descent (`_fit_coordinate_descent`,
sklearn's default), with sklearn's initialisations (`_initialize_nmf`: random, nndsvd, nndsvda, nndsvdar - the random numbers come from
the same numpy `RandomState(random_state)` streams, the truncated SVD is exact instead of randomised), its epsilon guards and its
stopping rules, so that the factors agree with sklearn's to rounding for the same start (tests/test_init_gpu.py).  All contractions go
through the library's GEMM (functional.gemm -> C ABI); element-wise steps are device tensor ops.  numpy in -> numpy out, tensor in ->
tensor out; there is no host path: without a CUDA device these functions raise.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import functional as F

EPSILON = float(np.finfo(np.float32).eps)          # sklearn.decomposition._nmf.EPSILON
_EPS64 = float(np.finfo(np.float64).eps)


def _device(device=None):
    if device is not None:
        device = torch.device(device)
    elif torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    if device is None or device.type != "cuda":
        raise RuntimeError("gpzoo_b200.initialisation runs on a CUDA device only (no host path)")
    return device


def _as_device(a, device=None, dtype=None):
    """-> (tensor on the device, was_numpy)."""
    if isinstance(a, torch.Tensor):
        dev = a.device if a.is_cuda else _device(device)
        return a.to(device=dev, dtype=dtype or a.dtype), False
    a = np.asarray(a)
    t = torch.from_numpy(np.ascontiguousarray(a)).to(_device(device))
    return (t.to(dtype) if dtype is not None else t), True


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _mm(A, B, ta=False, tb=False):
    """op(A) op(B) for 2-D operands on the library GEMM."""
    return F.gemm(A.contiguous().unsqueeze(0), B.contiguous().unsqueeze(0), ta=ta, tb=tb)[0]


# ------------------------------------------------------------------------------------------------
# small host-side formulas
# ------------------------------------------------------------------------------------------------
def lnormal_approx_dirichlet(L):
    """(mu, sigma) of the L independent lognormals that match the marginal mean and variance of a flat symmetric Dirichlet
    (utilities.py:237-250)."""
    sigma2 = math.log(2 * L) - math.log(L + 1)
    return -math.log(L) - sigma2 / 2.0, math.sqrt(sigma2)


def shrink_factors(Fm, shrinkage=0.2):
    """Rows pulled toward their mean, row sums preserved (utilities.py:301-306)."""
    a = shrinkage
    if 0 < a < 1:
        Fm = Fm * (1 - a) + a * Fm.sum(axis=1, keepdims=True) / float(Fm.shape[1])
    return Fm


def shrink_loadings(W, shrinkage=0.2):
    """Columns pulled toward their mean, column sums preserved (utilities.py:308-313)."""
    a = shrinkage
    if 0 < a < 1:
        W = W * (1 - a) + a * W.sum(axis=0) / float(W.shape[0])
    return W


def init_softplus(mat, minval=1e-5):
    """Inverse softplus log(exp(x) - 1 + minval) where x < 20, x itself above (utilities.py:38-44)."""
    t, was_np = _as_device(mat)
    out = torch.where(t < 20, torch.log(torch.expm1(t.clamp(max=20.0)) + minval), t)
    return _back(out, was_np)


def rescale_spatial_coords(X, box_side=4):
    """Centre the coordinates and scale them (aspect ratio kept) so that the bounding box has volume box_side^D
    (utilities.py:71-84).  Returns a new array; the reference also shifts / scales its argument in place."""
    t, was_np = _as_device(X)
    t = t - t.min(dim=0).values
    t = t * (box_side / torch.exp(torch.log(t.max(dim=0).values).mean()))
    return _back(t - t.mean(dim=0), was_np)


def build_group_distances(X, groupsX):
    """Distances between the groups' mean positions (utilities.py:14-24).  The reference takes `torch.mean(X[mask])` over BOTH
    coordinates, so every group is placed at (m, m) with m its scalar mean; that is what is reproduced."""
    X, was_np = _as_device(X)
    g, _ = _as_device(groupsX, device=X.device)
    n = int(g.max().item()) + 1
    onehot = torch.zeros(n, X.shape[0], dtype=X.dtype, device=X.device)
    onehot[g.long(), torch.arange(X.shape[0], device=X.device)] = 1.0
    m = _mm(onehot, X).sum(1) / (onehot.sum(1) * X.shape[1])
    pos = torch.stack((m, m), 1).float()
    return _back(F.cdist(pos, pos), was_np)


# ------------------------------------------------------------------------------------------------
# NMF
# ------------------------------------------------------------------------------------------------
def _truncated_svd(X, k):
    """Leading k singular triplets from the eigen-decomposition of the smaller Gram matrix (fp64)."""
    X64 = X.double()
    n, g = X64.shape
    if g <= n:
        w, V = torch.linalg.eigh(_mm(X64, X64, ta=True))            # G x G
        w, V = w.flip(0)[:k], V.flip(1)[:, :k]
        S = w.clamp_min(0).sqrt()
        U = _mm(X64, V) / S
        return U, S, V.t().contiguous()
    w, U = torch.linalg.eigh(_mm(X64, X64, tb=True))                # N x N
    w, U = w.flip(0)[:k], U.flip(1)[:, :k]
    S = w.clamp_min(0).sqrt()
    Vt = (_mm(U, X64, ta=True).t() / S).t().contiguous()
    return U, S, Vt


def initialize_nmf(X, n_components, init=None, eps=1e-6, random_state=None):
    """sklearn `_initialize_nmf` on the device -> (W0 N x L, H0 L x G).  Random draws come from numpy's
    `RandomState(random_state)` exactly as sklearn consumes them; NNDSVD uses an exact truncated SVD where sklearn uses a
    randomised one (the NNDSVD construction does not depend on the signs of the singular vectors)."""
    n, g = X.shape
    L = int(n_components)
    if init is not None and init != "random" and L > min(n, g):
        raise ValueError(f"init = '{init}' can only be used when n_components <= min(n_samples, n_features)")
    if init is None:
        init = "nndsvda" if L <= min(n, g) else "random"
    xmean = float(X.double().mean())
    if init == "random":
        avg = math.sqrt(xmean / L)
        rng = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
        H = np.abs(avg * rng.standard_normal(size=(L, g)))
        W = np.abs(avg * rng.standard_normal(size=(n, L)))
        return (torch.from_numpy(W).to(X.device, X.dtype), torch.from_numpy(H).to(X.device, X.dtype))
    if init not in ("nndsvd", "nndsvda", "nndsvdar"):
        raise ValueError(f"Invalid init parameter: got {init!r}")
    U, S, Vt = _truncated_svd(X, L)
    W = torch.zeros_like(U)
    H = torch.zeros_like(Vt)
    W[:, 0] = S[0].sqrt() * U[:, 0].abs()
    H[0] = S[0].sqrt() * Vt[0].abs()
    for j in range(1, L):
        x, y = U[:, j], Vt[j]
        xp, yp, xn, yn = x.clamp_min(0), y.clamp_min(0), (-x).clamp_min(0), (-y).clamp_min(0)
        mp, mn = xp.norm() * yp.norm(), xn.norm() * yn.norm()
        if float(mp) > float(mn):
            u, v, sigma = xp / xp.norm(), yp / yp.norm(), mp
        else:
            u, v, sigma = xn / xn.norm(), yn / yn.norm(), mn
        lbd = (S[j] * sigma).sqrt()
        W[:, j], H[j] = lbd * u, lbd * v
    W[W < eps] = 0
    H[H < eps] = 0
    if init == "nndsvda":
        W[W == 0] = xmean
        H[H == 0] = xmean
    elif init == "nndsvdar":
        rng = np.random.RandomState(random_state) if not isinstance(random_state, np.random.RandomState) else random_state
        zw, zh = (W == 0), (H == 0)
        W[zw] = torch.from_numpy(np.abs(xmean * rng.standard_normal(size=int(zw.sum())) / 100)).to(W)
        H[zh] = torch.from_numpy(np.abs(xmean * rng.standard_normal(size=int(zh.sum())) / 100)).to(H)
    return W.to(X.dtype), H.to(X.dtype)


def _beta_divergence(X, W, H, beta, square_root=True):
    """sklearn `_beta_divergence` for beta in {1, 2} (dense X)."""
    if beta == 2:
        res = float(((X - _mm(W, H)).double() ** 2).sum()) / 2.0
    else:
        nz = X > EPSILON                       # sklearn: only the entries with X > EPSILON enter sum X log(X / WH) and sum X
        Xn = X[nz].double()
        WHn = _mm(W, H)[nz].clamp_min(EPSILON).double()
        sum_WH = float((W.double().sum(0) * H.double().sum(1)).sum())
        res = float((Xn * torch.log(Xn / WHn)).sum()) + sum_WH - float(Xn.sum())
    res = max(res, 0.0)
    return math.sqrt(2 * res) if square_root else res


def _fit_mu(X, W, H, beta, max_iter, tol):
    """sklearn `_fit_multiplicative_update` (no regularisation, update_H=True), beta in {1 (Kullback-Leibler), 2 (Frobenius)}."""
    err0 = prev = _beta_divergence(X, W, H, beta)
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        if beta == 2:
            num = _mm(X, H, tb=True)
            den = _mm(W, _mm(H, H, tb=True))
        else:
            R = X / _mm(W, H).clamp_min(EPSILON)
            num = _mm(R, H, tb=True)
            den = H.sum(1)[None, :].expand_as(num).clone()
        den[den == 0] = EPSILON
        W = W * (num / den)
        if beta == 2:
            num = _mm(W, X, ta=True)
            den = _mm(_mm(W, W, ta=True), H)
        else:
            R = X / _mm(W, H).clamp_min(EPSILON)
            num = _mm(W, R, ta=True)
            ws = W.sum(0)
            ws[ws == 0] = 1.0
            den = ws[:, None].expand_as(num).clone()
        den[den == 0] = EPSILON
        H = H * (num / den)
        if beta <= 1:
            H[H < _EPS64] = 0.0
        if tol > 0 and n_iter % 10 == 0:
            err = _beta_divergence(X, W, H, beta)
            if (prev - err) / err0 < tol:
                break
            prev = err
    return W, H, n_iter


def _cd_sweep(W, HHt, XHt):
    """One pass of sklearn's `_update_cdnmf_fast`: components in order (Gauss-Seidel), all rows at once.  Returns the violation."""
    violation = 0.0
    L = W.shape[1]
    for t in range(L):
        grad = _mm(W, HHt[:, t:t + 1])[:, 0] - XHt[:, t]
        wt = W[:, t]
        pg = torch.where(wt == 0, grad.clamp_max(0), grad)
        violation += float(pg.abs().double().sum())
        hess = float(HHt[t, t])
        if hess != 0:
            W[:, t] = (wt - grad / hess).clamp_min(0)
    return violation


def _fit_cd(X, W, H, max_iter, tol):
    """sklearn `_fit_coordinate_descent` (Frobenius loss, no regularisation, shuffle=False)."""
    W = W.contiguous().clone()
    Ht = H.t().contiguous().clone()
    v0 = None
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        v = _cd_sweep(W, _mm(Ht, Ht, ta=True), _mm(X, Ht))
        v += _cd_sweep(Ht, _mm(W, W, ta=True), _mm(X, W, ta=True))
        if n_iter == 1:
            v0 = v
        if v0 == 0 or v / v0 <= tol:
            break
    return W, Ht.t().contiguous(), n_iter


def nmf(X, n_components, init=None, solver="cd", beta_loss="frobenius", tol=1e-4, max_iter=200, random_state=None, W=None, H=None,
        device=None, return_n_iter=False, **unsupported):
    """`sklearn.decomposition.NMF(n_components, init=..., solver=..., beta_loss=..., tol=..., max_iter=..., random_state=...)
    .fit_transform(X)` on the device -> (W N x L, H L x G) [, n_iter].  init='custom' takes W, H as the start."""
    bad = {k: v for k, v in unsupported.items() if k not in ("verbose", "shuffle") and v not in (None, 0, 0.0, False, "deprecated")}
    if bad:
        raise NotImplementedError(f"NMF options without a device implementation: {sorted(bad)}")
    Xd, was_np = _as_device(X, device)
    if not Xd.is_floating_point():
        Xd = Xd.double()
    if bool((Xd < 0).any()):
        raise ValueError("Negative values in data passed to NMF")
    beta = {"frobenius": 2, "kullback-leibler": 1, 2: 2, 1: 1, 2.0: 2, 1.0: 1}.get(beta_loss)
    if beta is None:
        raise NotImplementedError(f"beta_loss={beta_loss!r}: only 'frobenius' and 'kullback-leibler' run on the device")
    if solver == "cd" and beta != 2:
        raise ValueError("solver='cd' minimises the Frobenius loss only (as in sklearn)")
    if init == "custom":
        W0, _ = _as_device(W, Xd.device, Xd.dtype)
        H0, _ = _as_device(H, Xd.device, Xd.dtype)
        W0, H0 = W0.clone(), H0.clone()
    else:
        W0, H0 = initialize_nmf(Xd, n_components, init=init, random_state=random_state)
    if solver == "mu":
        Wf, Hf, n_iter = _fit_mu(Xd, W0, H0, beta, int(max_iter), float(tol))
    elif solver == "cd":
        Wf, Hf, n_iter = _fit_cd(Xd, W0, H0, int(max_iter), float(tol))
    else:
        raise ValueError(f"solver={solver!r}")
    out = (_back(Wf, was_np), _back(Hf, was_np))
    return out + (n_iter,) if return_n_iter else out


def regularized_nmf(Y, L, sz=1, pseudocount=1e-2, factors=None, loadings=None, shrinkage=0.2, **kwargs):
    """NMF of the (obs x feat) matrix Y, shrunk toward a flat symmetric Dirichlet; returns (log-scale factors N x L, non-negative
    loadings G x L) - utilities.py:253-299 with `sklearn.decomposition.NMF(L, **kwargs)` replaced by `nmf` above."""
    device = kwargs.pop("device", None)
    if factors is None or loadings is None:
        Yd, was_np = _as_device(Y, device)
        eF, H = nmf(Yd, L, **kwargs)
        W = H.t()
    else:
        eF, was_np = _as_device(factors, device)
        W, _ = _as_device(loadings, eF.device, eF.dtype)
    if not eF.is_floating_point():
        eF, W = eF.double(), W.double()
    szd = sz if isinstance(sz, (int, float)) else _as_device(sz, eF.device, eF.dtype)[0]
    W = shrink_loadings(W, shrinkage=shrinkage)
    wsum = W.sum(axis=0)
    eF = shrink_factors(eF * wsum, shrinkage=shrinkage)
    Fl = torch.log(pseudocount + eF) - (math.log(szd) if isinstance(szd, (int, float)) else torch.log(szd))
    prior_mu, _ = lnormal_approx_dirichlet(max(L, 1.1))
    wt_to_W = Fl.mean(axis=0) - prior_mu
    Fl = Fl - wt_to_W
    W = W * torch.exp(wt_to_W - torch.log(wsum))
    return _back(Fl, was_np), _back(W, was_np)


# ------------------------------------------------------------------------------------------------
# inducing points and the variational mean
# ------------------------------------------------------------------------------------------------
def kmeans_inducing(X, M, n_iter=50, seed=0, tol=1e-4):
    """M cluster centres of the coordinates (Lloyd's algorithm on the device; k-means++-style seeding by farthest-point sampling from
    a seeded start).  Stands in for `KMeans(n_clusters=M, random_state=..., n_init="auto").fit(X).cluster_centers_`
    (Slideseqv2_estimate_lengthscales.ipynb cell 16): same objective, not the same random stream.  -> (Z M x D, inertia)."""
    Xd, was_np = _as_device(X)
    Xd = Xd.float() if Xd.dtype not in (torch.float32, torch.float64) else Xd
    n = Xd.shape[0]
    if M > n:
        raise ValueError("more centres than points")
    g = torch.Generator(device="cpu").manual_seed(int(seed))
    first = int(torch.randint(n, (1,), generator=g))
    centres = torch.empty(M, Xd.shape[1], dtype=Xd.dtype, device=Xd.device)
    centres[0] = Xd[first]
    dmin = F.cdist(Xd, centres[:1])[:, 0]
    for j in range(1, M):                         # farthest-point seeding: one distance column per new centre
        centres[j] = Xd[int(dmin.argmax())]
        dmin = torch.minimum(dmin, F.cdist(Xd, centres[j:j + 1])[:, 0])
    inertia = prev = float("inf")
    for _ in range(n_iter):
        d = F.cdist(Xd, centres)                  # N x M by direct differences (C ABI)
        dm, assign = d.min(dim=1)
        inertia = float((dm.double() ** 2).sum())
        onehot = torch.zeros(M, n, dtype=Xd.dtype, device=Xd.device)
        onehot[assign, torch.arange(n, device=Xd.device)] = 1.0
        cnt = onehot.sum(1)
        new = _mm(onehot, Xd) / cnt.clamp_min(1)[:, None]
        centres = torch.where(cnt[:, None] > 0, new, centres)
        if prev - inertia <= tol * inertia:
            break
        prev = inertia
    return _back(centres, was_np), inertia


def project_to_inducing(kernel, Z, X, factors, jitter=1e-5):
    """Variational means that reproduce log-scale factors at the data: mu[l] = Kzz[l] (Kzx[l] Kxz[l] + j I)^-1 Kzx[l] f_l
    (Slideseqv2_estimate_lengthscales.ipynb cell 16: cholesky + cholesky_solve; NSF_Hybrid_benchmark.ipynb cell 11 uses a
    pseudo-inverse for the same product).  kernel: an L-factor gpzoo_b200 kernel; factors: L x N.  -> mu L x M."""
    Zd, was_np = _as_device(Z)
    par = next(iter(kernel.parameters()), None)
    if par is not None and par.is_cuda:              # compute in the kernel module's precision, on its device
        Zd = Zd.to(par.device, par.dtype)
    Xd, _ = _as_device(X, Zd.device, Zd.dtype)
    Fd, _ = _as_device(factors, Zd.device, Zd.dtype)
    with torch.no_grad():
        Kzx = kernel.forward(Zd, Xd)
        Kzz = kernel.forward(Zd, Zd)
        if Kzx.dim() == 2:
            Kzx, Kzz = Kzx.unsqueeze(0), Kzz.unsqueeze(0)
        Lf = Kzx.shape[0]
        A = F.gemm(Kzx.contiguous(), Kzx.contiguous(), tb=True)                       # L x M x M
        A.diagonal(dim1=-2, dim2=-1).add_(jitter)
        rhs = F.gemm(Kzx.contiguous(), Fd.reshape(Lf, -1, 1).contiguous())              # L x M x 1
        _, Ainv_chol = F.CholeskyInverse.apply(A)                                       # (chol, chol^-1)
        sol = F.gemm(Ainv_chol, F.gemm(Ainv_chol, rhs), ta=True)                        # A^-1 rhs = Li^T Li rhs
        mu = F.gemm(Kzz.contiguous(), sol)[:, :, 0]
    return _back(mu, was_np)
