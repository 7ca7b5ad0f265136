"""Deterministic synthetic problems shaped like the BASELINE.json configs.

Everything is drawn on the CPU from a seeded `torch.Generator` (so the oracle on the host and the
CUDA path on the device see bit-identical inputs) and then moved to the requested device.
No dataset is read; `data: "synthetic"` in bench.py refers to these generators.
"""
from __future__ import annotations

import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def grid_inducing(M, D, lo, hi, g, jitter_frac=0.15):
    """Jittered regular grid of ~M points in [lo,hi]^D (well separated -> well-conditioned Kzz)."""
    if D == 1:
        z = torch.linspace(lo, hi, M, dtype=torch.float64)[:, None]
    else:
        side = int(math.ceil(M ** (1.0 / D)))
        ax = torch.linspace(lo, hi, side, dtype=torch.float64)
        z = torch.cartesian_prod(*([ax] * D))[:M]
    step = (hi - lo) / max(2, int(round(M ** (1.0 / D))))
    z = z + jitter_frac * step * (2 * torch.rand(z.shape, generator=g, dtype=torch.float64) - 1)
    return z


def nsf_problem(N=256, M=36, L=3, G=16, D=2, E=1, seed=0, coord_scale=2.0, lengthscale=None,
                sigma=1.0, jitter=1e-2, z_mode="grid", lu_scale=0.05, mu_scale=0.5,
                dtype=torch.float64, device="cpu", n_groups=0):
    """Inputs and parameters of an NSF2(SVGP(NSF_RBF)) model (config 2 shape when called with
    N=32768, M=1024, L=10, G=2000, coord_scale=100, lengthscale=1.7, jitter=0.1)."""
    g = _gen(seed)
    X = coord_scale * (2 * torch.rand(N, D, generator=g, dtype=torch.float64) - 1)
    if z_mode == "grid":
        Z = grid_inducing(M, D, -coord_scale, coord_scale, g)
    else:  # random subset of X, as Slideseq_NSF_newest_version.ipynb:362-368 does
        Z = X[torch.randperm(N, generator=g)[:M]].clone()
    if lengthscale is None:
        lengthscale = 1.2 * 2 * coord_scale / max(2.0, M ** (1.0 / D))
    ls = lengthscale * (1 + 0.1 * torch.arange(L, dtype=torch.float64) / max(1, L)).reshape(L, 1, 1)
    sg = sigma * (1 + 0.05 * torch.arange(L, dtype=torch.float64) / max(1, L)).reshape(L, 1, 1)
    mu = mu_scale * torch.randn(L, M, generator=g, dtype=torch.float64)
    Lu_raw = lu_scale * torch.randn(L, M, M, generator=g, dtype=torch.float64)
    W = torch.rand(G, L, generator=g, dtype=torch.float64)
    V = 1.0 + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    # smooth synthetic rate field -> sparse Poisson counts (mean ~0.3, Slide-seq-like sparsity)
    freq = math.pi / coord_scale
    basis = torch.stack([torch.sin(freq * (l + 1) * X[:, 0] * 0.5) * torch.cos(freq * (l + 1) * X[:, -1] * 0.5)
                         for l in range(L)])                                    # L x N
    rate = 0.3 * torch.nn.functional.softplus(W) @ torch.exp(0.5 * basis) / L    # G x N
    y = torch.poisson(rate, generator=g)
    eps = torch.randn(E, L, N, generator=g, dtype=torch.float64)
    out = dict(X=X, Z=Z, sigma=sg, lengthscale=ls, mu=mu, Lu_raw=Lu_raw, W=W, V=V, y=y, eps=eps)
    if n_groups:
        # spatially clustered group labels (config 4): nearest of n_groups random centres
        centres = coord_scale * (2 * torch.rand(n_groups, D, generator=g, dtype=torch.float64) - 1)
        out["groupsX"] = torch.cdist(X, centres).argmin(1)
        out["groupsZ"] = torch.cdist(Z, centres).argmin(1)
        out["gdp"] = (0.7 + 0.1 * torch.arange(L, dtype=torch.float64)).reshape(L, 1, 1)
        gd = torch.rand(n_groups, n_groups, generator=g, dtype=torch.float64) + 0.5
        gd = 0.5 * (gd + gd.t())
        gd.fill_diagonal_(0.0)
        out["group_distances"] = gd
    out = {k: (v.to(dtype) if v.is_floating_point() else v).to(device) for k, v in out.items()}
    out["jitter"] = float(jitter)
    return out


def hash_uniform(*shape, salt=0):
    """Reproducible uniform(-1, 1) array from integer arithmetic only (no RNG stream, no floating-point intermediate), so a
    fixture too large to commit (the 10 x 1024 x 1024 raw Lu of the config-2-conditioned golden) is regenerated bit-exactly on
    any machine: three rounds of a 31-bit LCG + xor-shift over sum_d idx_d * prime_d + salt, top 24 bits mapped to [-1, 1)."""
    primes = (73856093, 19349663, 83492791, 49979687, 86028121)
    m31 = 1 << 31                                         # every product below stays under 2^62: no int64 wrap-around
    h = torch.zeros(shape, dtype=torch.int64) + (int(salt) * 1013904223) % m31
    for d, n in enumerate(shape):
        view = [1] * len(shape)
        view[d] = n
        h = (h + (torch.arange(n, dtype=torch.int64) * primes[d % len(primes)]).view(view)) % m31
    for mult in (1103515245, 1664525, 22695477):
        h = (h * mult + 12345) % m31
        h = h ^ (h >> 15)
    return (h >> 7).to(torch.float64) / float(1 << 23) - 1.0


def regression_problem(N=2000, M=100, D=2, E=20, seed=0, dtype=torch.float64, device="cpu", jitter=1e-3):
    """Config 1: SVGP regression, y = 2 sin(2 x0) cos(x1) + N(0, 0.1^2), X ~ U(-5,5)^D."""
    g = _gen(seed)
    X = 5.0 * (2 * torch.rand(N, D, generator=g, dtype=torch.float64) - 1)
    y = 2 * torch.sin(2 * X[:, 0]) * torch.cos(X[:, -1]) + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    Z = grid_inducing(M, D, -5.0, 5.0, g)
    mu = 0.3 * torch.randn(M, generator=g, dtype=torch.float64)
    Lu_raw = 0.05 * torch.randn(M, M, generator=g, dtype=torch.float64)
    eps = torch.randn(E, N, generator=g, dtype=torch.float64)
    out = dict(X=X, y=y, Z=Z, mu=mu, Lu_raw=Lu_raw, eps=eps,
               sigma=torch.tensor(1.0, dtype=torch.float64), lengthscale=torch.tensor(1.0, dtype=torch.float64),
               noise=torch.tensor(0.1, dtype=torch.float64))
    out = {k: v.to(dtype).to(device) for k, v in out.items()}
    out["jitter"] = float(jitter)
    return out
