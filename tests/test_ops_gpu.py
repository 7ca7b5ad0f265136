"""GPU: each CUDA kernel (through the C ABI) against the CPU oracle / plain fp64 torch on seeded inputs."""
import math

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tol(dt):
    return 1e-11 if dt == torch.float64 else 2e-5


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_layouts(dt, ta, tb):
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(0)
    b, m, n, k = 3, 150, 77, 203
    A = torch.randn((b, k, m) if ta else (b, m, k), generator=g, dtype=torch.float64)
    B = torch.randn((b, n, k) if tb else (b, k, n), generator=g, dtype=torch.float64)
    ref = (A.transpose(1, 2) if ta else A) @ (B.transpose(1, 2) if tb else B)
    out = F.gemm(A.to(DEV, dt), B.to(DEV, dt), ta=ta, tb=tb)
    assert relerr(out, ref) < _tol(dt)
    out2 = F.gemm(A.to(DEV, dt), B.to(DEV, dt), ta=ta, tb=tb, splitk=4)
    assert relerr(out2, ref) < _tol(dt)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_gemm_triangular_flags(dt):
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(1)
    b, m = 2, 300
    Lo = torch.tril(torch.randn(b, m, m, generator=g, dtype=torch.float64))
    R = torch.randn(b, m, 190, generator=g, dtype=torch.float64)
    assert relerr(F.gemm(Lo.to(DEV, dt), R.to(DEV, dt), a_tri=1), Lo @ R) < _tol(dt)
    assert relerr(F.gemm(Lo.to(DEV, dt), R.to(DEV, dt), ta=True, a_tri=2), Lo.transpose(1, 2) @ R) < _tol(dt)
    Lo2 = torch.tril(torch.randn(b, m, m, generator=g, dtype=torch.float64))
    assert relerr(F.gemm(Lo.to(DEV, dt), Lo2.to(DEV, dt), a_tri=1, b_tri=1, d_tri=1), Lo @ Lo2) < _tol(dt)
    out = F.gemm(R.to(DEV, dt), R.to(DEV, dt), tb=True, d_tri=1, splitk=2)
    assert relerr(out, torch.tril(R @ R.transpose(1, 2))) < _tol(dt)
    up = F.gemm(Lo.to(DEV, dt), Lo2.to(DEV, dt), ta=True, tb=True, a_tri=2, b_tri=2)
    assert relerr(up, Lo.transpose(1, 2) @ Lo2.transpose(1, 2)) < _tol(dt)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("M", [25, 64, 200, 515, 1100, 1600, 2048])      # > 1536: tensor-core divide and conquer in fp32
def test_cholesky_and_inverse(dt, M):
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(M)
    L = 3
    Q = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    K = Q @ Q.transpose(1, 2) / M + torch.eye(M, dtype=torch.float64)
    Lc, Linv = F.CholeskyInverse.apply(K.to(DEV, dt))
    ref = torch.linalg.cholesky(K)
    tol = 1e-11 if dt == torch.float64 else 5e-5
    assert relerr(Lc, ref) < tol
    assert relerr(Linv, torch.linalg.inv(ref)) < tol
    assert torch.equal(Lc.triu(1), torch.zeros_like(Lc)) and torch.equal(Linv.triu(1), torch.zeros_like(Linv))


def test_cholesky_not_pd_raises():
    from gpzoo_b200 import functional as F
    K = torch.eye(40, dtype=torch.float64).repeat(2, 1, 1)
    K[1, 17, 17] = -1.0
    with pytest.raises(torch.linalg.LinAlgError):
        F.CholeskyInverse.apply(K.to(DEV))


def test_cholesky_backward_fp64():
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(7)
    L, M = 2, 90
    Q = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    K = (Q @ Q.transpose(1, 2) / M + torch.eye(M, dtype=torch.float64))
    w1 = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    w2 = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    Kc = K.clone().requires_grad_(True)
    Lr = torch.linalg.cholesky(Kc)
    ((Lr * w1).sum() + (torch.linalg.inv(Lr) * w2.tril()).sum()).backward()
    Kg = K.to(DEV).requires_grad_(True)
    Lc, Linv = F.CholeskyInverse.apply(Kg)
    ((Lc * w1.to(DEV)).sum() + (Linv * w2.tril().to(DEV)).sum()).backward()
    sym = lambda t: 0.5 * (t + t.transpose(1, 2))
    assert relerr(sym(Kg.grad), sym(Kc.grad)) < 1e-10


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_kernel_build_vs_oracle(dt):
    from oracle import gpzoo_oracle as O
    import gpzoo_b200 as gz
    z = {k: torch.from_numpy(v) for k, v in np.load(GOLDEN + "/kernels.npz").items()}
    X, Z, gX, gZ = z["X"], z["Z"], z["gX"], z["gZ"]
    tol = 1e-12 if dt == torch.float64 else 3e-6
    k = gz.kernels.RBF(sigma=1.3, lengthscale=0.7).to(DEV)
    assert relerr(k(X.to(DEV, dt), Z.to(DEV, dt)), z["rbf"]) < max(tol, 1e-7)     # ctor rounds params to fp32
    nk = gz.kernels.NSF_RBF(sigma=1.1, lengthscale=0.9, L=3).to(DEV).to(dt)
    with torch.no_grad():
        nk.lengthscale.copy_(z["nsf_rbf_ls"].to(DEV, dt))
        nk.sigma.copy_((1.1 * torch.ones(3, 1, 1)).double().to(DEV, dt))
    assert relerr(nk(X.to(DEV, dt), Z.to(DEV, dt)), z["nsf_rbf"]) < tol
    assert relerr(nk(X.to(DEV, dt), X.to(DEV, dt), diag=True), O.rbf_diag(X, nk.sigma.detach().cpu().double())) < tol
    mk = gz.kernels.MGGP_NSF_RBF(sigma=1.2, lengthscale=0.8, group_diff_param=1.3, n_groups=4, L=2).to(DEV).double().to(dt)
    mk.embedding = torch.nn.Parameter(z["mggp_embedding"].to(DEV, dt), requires_grad=False)
    out = mk(X.to(DEV, dt), Z.to(DEV, dt), gX.to(DEV), gZ.to(DEV))
    assert relerr(out, z["mggp_nsf_rbf"]) < max(tol, 1e-7)
    mr = gz.kernels.MGGP_RBF(sigma=1.2, lengthscale=0.8, group_diff_param=1.7, n_groups=4).to(DEV)
    mr.embedding = z["mggp_embedding"].to(DEV, dt)
    assert relerr(mr(X.to(DEV, dt), Z.to(DEV, dt), gX.to(DEV), gZ.to(DEV)), z["mggp_rbf"]) < max(tol, 1e-7)
    # default embedding (ones - eye): same-group r^2 = 0, cross-group ~1.000002
    e = gz.kernels.embed_distance_matrix(torch.ones(4, 4) - torch.eye(4))
    assert relerr(O.squared_dist(e, e), O.squared_dist(z["mggp_default_embedding"].float(), z["mggp_default_embedding"].float())) < 1e-5
    # return_distance
    _, d = k(X.to(DEV, dt), Z.to(DEV, dt), return_distance=True)
    assert relerr(d, torch.cdist(X, Z, compute_mode="donot_use_mm_for_euclid_dist")) < tol


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("mg", [False, True])
def test_kernel_build_backward(dt, mg):
    """Gradients w.r.t. both point sets and all hyper-parameters vs autograd of the oracle (ragged sizes)."""
    from oracle import gpzoo_oracle as O
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(3)
    n1, n2, L, ng = 37, 1030, 3, 4
    x1 = torch.randn(n1, 2, generator=g, dtype=torch.float64)
    x2 = torch.randn(n2, 2, generator=g, dtype=torch.float64)
    sg = 1 + 0.1 * torch.rand(L, generator=g, dtype=torch.float64)
    ls = 0.8 + 0.4 * torch.rand(L, generator=g, dtype=torch.float64)
    gdp = 0.5 + torch.rand(L, generator=g, dtype=torch.float64)
    emb = torch.randn(ng, ng, generator=g, dtype=torch.float64)
    g1 = torch.randint(0, ng, (n1,), generator=g)
    g2 = torch.randint(0, ng, (n2,), generator=g)
    Wt = torch.randn(L, n1, n2, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (x1, x2, sg, ls, gdp)]
    a, b, s, l, gd = leaves
    if mg:
        Kref = O.mggp_nsf_rbf(a, b, g1, g2, s.reshape(L, 1, 1), l.reshape(L, 1, 1), gd.reshape(L, 1, 1), emb)
    else:
        Kref = O.nsf_rbf(a, b, s.reshape(L, 1, 1), l.reshape(L, 1, 1))
    (Kref * Wt).sum().backward()
    dl = [t.detach().to(DEV, dt).requires_grad_(True) for t in (x1, x2, sg, ls, gdp)]
    a2, b2, s2, l2, gd2 = dl
    if mg:
        d = emb[:, None, :] - emb[None, :, :]
        r2 = (d * d).sum(-1).to(DEV, dt)
        K = F.KernelBuild.apply(a2, b2, s2, l2, gd2 ** 2, r2, g1.to(DEV), g2.to(DEV), 1.0, 0.0)
    else:
        K = F.KernelBuild.apply(a2, b2, s2, l2, None, None, None, None, 1.0, 0.0)
    tol = 1e-11 if dt == torch.float64 else 2e-5
    assert relerr(K, Kref) < tol
    (K * Wt.to(DEV, dt)).sum().backward()
    for i, (r, c) in enumerate(zip(leaves, dl)):
        if i == 4 and not mg:
            continue
        assert relerr(c.grad, r.grad) < tol, i


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_predict_and_kl_vs_oracle(dt):
    """Whiten + Predict + MvnKL against the reference formulation (cholesky_solve, W(S-Kzz)W, torch KL)."""
    from oracle import gpzoo_oracle as O
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(5)
    L, M, N = 2, 70, 333
    Z = torch.rand(M, 2, generator=g, dtype=torch.float64) * 4
    X = torch.rand(N, 2, generator=g, dtype=torch.float64) * 4
    sg = torch.ones(L, 1, 1, dtype=torch.float64)
    ls = torch.tensor([0.6, 0.9], dtype=torch.float64).reshape(L, 1, 1)
    mu = torch.randn(L, M, generator=g, dtype=torch.float64)
    Lu_raw = 0.1 * torch.randn(L, M, M, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (mu, Lu_raw)]
    Kzx, Kzz, Kxx = O.nsf_rbf(Z, X, sg, ls), O.nsf_rbf(Z, Z, sg, ls), O.rbf_diag(X, sg)
    Kzx_r, Kzz_r = Kzx.clone().requires_grad_(True), Kzz.clone().requires_grad_(True)
    mean, var, Lu, Lc = O.svgp(Kxx, Kzx_r, Kzz_r, leaves[0], leaves[1], 1e-2, -1e30)
    kl = O.mvn_kl(leaves[0], Lu, Lc)
    wm = torch.randn(L, N, generator=g, dtype=torch.float64)
    wv = torch.randn(L, N, generator=g, dtype=torch.float64)
    ((mean * wm).sum() + (var * wv).sum() - kl.sum()).backward()

    dev = lambda t: t.detach().to(DEV, dt)
    mu_d, Lur_d = dev(mu).requires_grad_(True), dev(Lu_raw).requires_grad_(True)
    Kzx_d = dev(Kzx).requires_grad_(True)
    Kzz_d = dev(Kzz + 1e-2 * torch.eye(M, dtype=torch.float64)).requires_grad_(True)
    Lc_d, Linv_d = F.CholeskyInverse.apply(Kzz_d)
    Lu_d = F.LowerCholesky.apply(Lur_d)
    T_d, q_d = F.Whiten.apply(Linv_d, Lu_d, mu_d)
    mean_d, var_d = F.Predict.apply(dev(Kxx.contiguous()), Kzx_d, Linv_d, T_d, q_d)
    kl_d = F.MvnKL.apply(T_d, q_d, Lc_d, Lu_d)
    tol = 1e-10 if dt == torch.float64 else 1e-4
    assert relerr(mean_d, mean) < tol and relerr(var_d, var) < tol and relerr(kl_d, kl) < tol
    ((mean_d * dev(wm)).sum() + (var_d * dev(wv)).sum() - kl_d.sum()).backward()
    sym = lambda t: 0.5 * (t + t.transpose(1, 2))
    assert relerr(mu_d.grad, leaves[0].grad) < tol
    assert relerr(Lur_d.grad, leaves[1].grad) < tol
    assert relerr(Kzx_d.grad, Kzx_r.grad) < tol
    assert relerr(sym(Kzz_d.grad), sym(Kzz_r.grad)) < tol


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("case", ["full", "idx", "hybrid_raw"])
def test_poisson_fused_vs_oracle(dt, case):
    from oracle import gpzoo_oracle as O
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(9)
    G, Fn, Ntot, E = 45, 5, 300, 3
    nvar = 3
    y = torch.poisson(torch.rand(G, Ntot, generator=g, dtype=torch.float64) * 2, generator=g)
    idx = torch.randperm(Ntot, generator=g)[:170] if case != "full" else None
    B = Ntot if idx is None else len(idx)
    W = torch.rand(G, Fn, generator=g, dtype=torch.float64)
    V = 1 + 0.2 * torch.randn(Ntot, generator=g, dtype=torch.float64)
    mean = 0.3 * torch.randn(Fn, B, generator=g, dtype=torch.float64)
    spread = 0.05 + 0.3 * torch.rand(Fn, B, generator=g, dtype=torch.float64)
    spread[0, :7] = 1e-3                                   # exercise the variance clamp (5e-2)
    eps = torch.randn(E, Fn, B, generator=g, dtype=torch.float64)
    soft = case != "hybrid_raw"
    leaves = [t.clone().requires_grad_(True) for t in (W, V, mean, spread)]
    Wl, Vl, ml, sl = leaves
    sd = torch.cat((torch.clamp(sl[:nvar], min=5e-2).sqrt(), sl[nvar:]), 0)
    Fs = ml + eps * sd
    rate = O.poisson_rate(Wl, Fs, Vl if idx is None else Vl[idx], softplus_W=soft)
    ll = O.poisson_loglik(y if idx is None else y[:, idx], rate, with_lgamma=True)
    ll.backward()
    dl = [t.detach().to(DEV, dt).requires_grad_(True) for t in (W, V, mean, spread)]
    out = F.PoissonLL.apply(y.to(DEV, dt), None if idx is None else idx.to(DEV), dl[0], dl[1], dl[2], dl[3], eps.to(DEV, dt),
                            nvar, 5e-2, soft, True)
    out.backward()
    tol = 1e-11 if dt == torch.float64 else 3e-5
    assert relerr(out, ll) < tol
    for r, c in zip(leaves, dl):
        assert relerr(c.grad, r.grad) < tol
    # compatibility path: materialised rate and its backward
    dl2 = [t.detach().to(DEV, dt).requires_grad_(True) for t in (W, V)]
    Fd = Fs.detach().to(DEV, dt).requires_grad_(True)
    r2 = F.poisson_rate(dl2[0], dl2[1], None if idx is None else idx.to(DEV), Fd, soft)
    assert relerr(r2, rate) < tol


@pytest.mark.parametrize("ydt", [torch.uint8, torch.int16, torch.int32])
@pytest.mark.parametrize("case", ["full", "idx", "ragged"])
def test_poisson_integer_counts(ydt, case):
    """Counts stored as uint8 / int16 / int32 are read as stored (gpz_poisson_fwdbwd_yt_f32): the same numbers as the float-y call."""
    from gpzoo_b200 import _cabi, functional as F
    g = torch.Generator().manual_seed(21)
    G, Fn, E = 77, 6, 2
    Ntot = {"full": 512, "idx": 400, "ragged": 333}[case]            # 333: odd row length -> the scalar-load path
    y = torch.poisson(torch.rand(G, Ntot, generator=g) * 3, generator=g)
    y[3, 5] = 200.0
    idx = torch.randperm(Ntot, generator=g)[:257].to(DEV) if case == "idx" else None
    B = Ntot if idx is None else 257
    W, V = torch.rand(G, Fn, generator=g), 1 + 0.2 * torch.randn(Ntot, generator=g)
    mean, spread = 0.3 * torch.randn(Fn, B, generator=g), 0.05 + 0.3 * torch.rand(Fn, B, generator=g)
    eps = torch.randn(E, Fn, B, generator=g)

    def run(yy):
        lv = [t.clone().to(DEV).requires_grad_(True) for t in (W, V, mean, spread)]
        _cabi.profile = {}
        out = F.PoissonLL.apply(yy.to(DEV), idx, lv[0], lv[1], lv[2], lv[3], eps.to(DEV), 3, 5e-2, True, True)
        calls, _cabi.profile = set(_cabi.profile), None
        out.backward()
        return out, [t.grad for t in lv], calls
    o_f, g_f, c_f = run(y)
    o_i, g_i, c_i = run(y.to(ydt))
    assert "poisson_fwdbwd" in c_f and "poisson_fwdbwd_yt" in c_i
    assert relerr(o_i, o_f) < 1e-7                        # same arithmetic; the D2 atomics over the gene ranges are not ordered
    for a, b in zip(g_i, g_f):
        assert relerr(a, b) < 1e-6
    # and fp64 parameters with integer counts: converted on the device
    lv = [t.double().to(DEV).requires_grad_(True) for t in (W, V, mean, spread)]
    o64 = F.PoissonLL.apply(y.to(ydt).to(DEV), idx, lv[0], lv[1], lv[2], lv[3], eps.double().to(DEV), 3, 5e-2, True, True)
    assert relerr(o_i, o64) < 1e-6


@pytest.mark.parametrize("bk", [0, 1])
def test_umma_gemm_split_tf32(bk):
    """tcgen05/TMA split-TF32 GEMM vs fp64: ragged sizes, batch, triangular skipping, Cin, lo output, split-K."""
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(11)
    b, m, n, k = 2, 300, 520, 200
    A = torch.randn(b, m, k, generator=g)
    B = torch.randn((b, n, k) if bk else (b, k, n), generator=g)
    ref = A.double() @ (B.double().transpose(1, 2) if bk else B.double())
    out, lo = F.umma_gemm(A.to(DEV), B.to(DEV), bk, want_lo=True)
    assert relerr(out, ref) < 5e-6
    assert relerr(out.double() - lo.double(), (out.view(torch.int32) & -8192).view(torch.float32)) < 1e-12
    one = F.umma_gemm(A.to(DEV), B.to(DEV), bk, n_terms=1)                  # plain TF32: only ~1e-3
    assert 1e-5 < relerr(one, ref) < 3e-3
    Cin = torch.randn(b, m, n, generator=g)
    out2 = F.umma_gemm(A.to(DEV), B.to(DEV), bk, Cin=Cin.to(DEV), alpha=0.5)
    assert relerr(out2, 0.5 * ref + Cin.double()) < 5e-6
    out3 = F.umma_gemm(A.to(DEV), B.to(DEV), bk, splitk=3)
    assert relerr(out3, ref) < 5e-6
    # square, triangular A (lower) and lower-triangular output
    m2 = 384
    Lo = torch.tril(torch.randn(b, m2, m2, generator=g))
    R = torch.randn((b, 640, m2) if bk else (b, m2, 640), generator=g)
    refL = Lo.double() @ (R.double().transpose(1, 2) if bk else R.double())
    assert relerr(F.umma_gemm(Lo.to(DEV), R.to(DEV), bk, a_tri=1), refL) < 5e-6
    Up = Lo.transpose(1, 2).contiguous()
    refU = Up.double() @ (R.double().transpose(1, 2) if bk else R.double())
    assert relerr(F.umma_gemm(Up.to(DEV), R.to(DEV), bk, a_tri=2), refU) < 5e-6
    if bk:
        S = torch.randn(b, m2, 1000, generator=g)
        refS = torch.tril(S.double() @ S.double().transpose(1, 2))
        assert relerr(F.umma_gemm(S.to(DEV), S.to(DEV), 1, d_tri=1, splitk=2), refS) < 2e-5


@pytest.mark.parametrize("bk", [0, 1])
def test_umma_gemm_split_fp16(bk):
    """tcgen05/TMA split-FP16 GEMM vs fp64: ragged sizes, batch, badly scaled operands, triangular skipping, split-K, fp16 plane
    output with its own scale, max |D| tracking."""
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(12)
    b, m, n, k = 2, 296, 520, 200
    A = torch.randn(b, m, k, generator=g) * 1e3                      # scales are per batch entry and exact powers of two
    B = torch.randn((b, n, k) if bk else (b, k, n), generator=g) * torch.tensor([1e-4, 30.0])[:, None, None]
    ref = A.double() @ (B.double().transpose(1, 2) if bk else B.double())
    Ap, Bp = F.split16(A.to(DEV)), F.split16(B.to(DEV))
    assert relerr(Ap[0].double() + Ap[1].double(), A.double() * Ap[2].cpu().double()[:, None, None]) < 1e-6
    assert float(Ap[0].float().abs().max()) <= 2.0 ** 15
    out = F.umma_gemm16(Ap, Bp, bk)
    assert relerr(out[0], ref[0]) < 5e-6 and relerr(out[1], ref[1]) < 5e-6
    one = F.umma_gemm16(Ap, Bp, bk, n_terms=1)                        # plain fp16 operands: only ~3e-4
    assert 1e-5 < relerr(one[0], ref[0]) < 3e-3
    assert relerr(F.umma_gemm16(Ap, Bp, bk, splitk=3)[1], ref[1]) < 5e-6
    sd = torch.tensor([2.0 ** 3, 2.0 ** -9], device=DEV)
    (Dh, Dl), amax = F.umma_gemm16(Ap, Bp, bk, out_planes=True, out_scale=sd, want_amax=True)
    rec = (Dh.double() + Dl.double()) / sd.double()[:, None, None]
    assert relerr(rec[0], ref[0]) < 5e-6 and relerr(rec[1], ref[1]) < 5e-6
    assert relerr(amax, ref.abs().amax((1, 2))) < 1e-5
    m2 = 384
    Lo = torch.tril(torch.randn(b, m2, m2, generator=g))
    R = torch.randn((b, 640, m2) if bk else (b, m2, 640), generator=g)
    refL = Lo.double() @ (R.double().transpose(1, 2) if bk else R.double())
    h, l, hT, lT, sc = F.split16(Lo.to(DEV), transpose=True)
    Rp = F.split16(R.to(DEV))
    assert relerr(F.umma_gemm16((h, l, sc), Rp, bk, a_tri=1), refL) < 5e-6
    refU = Lo.double().transpose(1, 2) @ (R.double().transpose(1, 2) if bk else R.double())
    assert relerr(F.umma_gemm16((hT, lT, sc), Rp, bk, a_tri=2), refU) < 5e-6
    if bk:
        S = torch.randn(b, m2, 1000, generator=g)
        Sp = F.split16(S.to(DEV))
        refS = torch.tril(S.double() @ S.double().transpose(1, 2))
        assert relerr(F.umma_gemm16(Sp, Sp, 1, d_tri=1, splitk=2), refS) < 2e-5


@pytest.mark.parametrize("arith", ["fp16x3", "tf32x3"])
def test_predict_tensor_core_path_vs_exact(arith):
    """fp32 tensor-core predict (split-FP16 planes or split-TF32) against the fp64 CUDA-core path on the same inputs, fwd and
    bwd, including the kernel build that writes the operand planes."""
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(13)
    L, M, N = 2, 192, 768
    Z = torch.rand(M, 2, generator=g, dtype=torch.float64) * 10
    X = torch.rand(N, 2, generator=g, dtype=torch.float64) * 10
    sg = torch.tensor([1.0, 1.1], dtype=torch.float64)
    ls = torch.tensor([0.7, 1.0], dtype=torch.float64)
    mu = torch.randn(L, M, generator=g, dtype=torch.float64)
    Lur = 0.1 * torch.randn(L, M, M, generator=g, dtype=torch.float64)
    wm = torch.randn(L, N, generator=g, dtype=torch.float64)
    wv = torch.randn(L, N, generator=g, dtype=torch.float64)
    res = {}
    for dt in (torch.float64, torch.float32):
        d = lambda t: t.to(DEV, dt)
        Zd, mud, Lud = d(Z).requires_grad_(True), d(mu).requires_grad_(True), d(Lur).requires_grad_(True)
        lsd, sgd = d(ls).requires_grad_(True), d(sg).requires_grad_(True)
        Kzz = F.KernelBuild.apply(Zd, Zd, sgd, lsd, None, None, None, None, 1.0, 0.05)
        Lc, Linv = F.CholeskyInverse.apply(Kzz)
        Lu = F.LowerCholesky.apply(Lud)
        T, q = F.Whiten.apply(Linv, Lu, mud)
        Kxx = (sgd ** 2)[:, None].expand(-1, N).contiguous()
        if dt == torch.float32 and arith == "fp16x3":
            Kzx, Kh, Kl, sK = F.KernelBuildH.apply(Zd, d(X), sgd, lsd, None, None, None, None, 1.0, 0.0)
            mean, var = F.PredictH.apply(Kxx, Kzx, Linv, T, q, Kh, Kl, sK)
        else:
            out = F.KernelBuild.apply(Zd, d(X), sgd, lsd, None, None, None, None, 1.0, 0.0, dt == torch.float32)
            Kzx, Kzx_lo = out if isinstance(out, tuple) else (out, None)
            mean, var = F.Predict.apply(Kxx, Kzx, Linv, T, q, Kzx_lo)
        ((mean * d(wm)).sum() + (var * d(wv)).sum()).backward()
        res[dt] = dict(mean=mean, var=var, gZ=Zd.grad, gmu=mud.grad, gLu=Lud.grad, gls=lsd.grad, gsg=sgd.grad)
    for k in res[torch.float64]:
        assert relerr(res[torch.float32][k], res[torch.float64][k]) < 1e-4, k


def test_split_fp16_overflow_guard():
    """The fp16 planes rely on |A[:,n]|^2 <= Kxx[n]; if a caller breaks that contract (here: a Kxx that is far too small for the
    kernel matrix it is paired with) the planes overflow and the deferred check must say so instead of returning garbage."""
    from gpzoo_b200 import _cabi, functional as F
    g = torch.Generator().manual_seed(14)
    L, M, N = 1, 128, 512
    Z = (torch.rand(M, 2, generator=g) * 10).to(DEV)
    X = (torch.rand(N, 2, generator=g) * 10).to(DEV)
    sg, ls = torch.ones(L, device=DEV), torch.ones(L, device=DEV)
    Kzz = F.KernelBuild.apply(Z, Z, sg, ls, None, None, None, None, 1.0, 0.05)
    Lc, Linv = F.CholeskyInverse.apply(Kzz)
    T, q = F.Whiten.apply(Linv, torch.eye(M, device=DEV).expand(L, -1, -1).contiguous(), torch.zeros(L, M, device=DEV))
    Kzx, Kh, Kl, sK = F.KernelBuildH.apply(Z, X, sg, ls, None, None, None, None, 1.0, 0.0)
    ok = (sg ** 2)[:, None].expand(-1, N).contiguous()
    F.PredictH.apply(ok, Kzx, Linv, T, q, Kh, Kl, sK)                      # the contract holds: no complaint
    with pytest.raises(_cabi.GpzError, match="fp16 range"):
        F.PredictH.apply(ok * 1e-12, Kzx, Linv, T, q, Kh, Kl, sK)


@pytest.mark.parametrize("dt", [torch.float32, torch.float64])
def test_fused_adam_matches_torch(dt):
    """gpzoo_b200.optim.Adam (one multi-tensor launch, clamp fused) against torch.optim.Adam + clamp_ (utilities.py:621-623)."""
    import gpzoo_b200 as gz
    g = torch.Generator().manual_seed(21)
    shapes = [(7,), (33, 5), (2, 129, 65), (1,), (3000,)]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g, dtype=dt).to(DEV)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=3e-2, betas=(0.9, 0.99), eps=1e-8)
    o_ours = gz.optim.Adam(ours, lr=3e-2, betas=(0.9, 0.99), eps=1e-8, clamp_nonneg=[ours[1]])
    for it in range(6):
        for a, b in zip(ref, ours):
            gr = torch.randn(a.shape, generator=g, dtype=dt).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        if it == 3:
            ref[0].grad = ours[0].grad = None                     # a parameter without gradient is skipped (its step count too)
        o_ref.step()
        with torch.no_grad():
            ref[1].clamp_(min=0)
        o_ours.step()
    tol = 1e-6 if dt == torch.float32 else 1e-13
    for a, b in zip(ref, ours):
        assert relerr(b, a) < tol
    assert float(ours[1].min()) >= 0.0


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_matern32_kernel_vs_oracle(dt):
    """batched_Matern32 (kernels.py:6-30) through the fused kernel build (kind = 1): forward against the
    golden vector the unmodified reference produced, backward against autograd of the oracle restatement on distinct points, and
    finite, correct gradients on coincident points (Kzz), where the reference's autograd of sqrt(sum diff^2) yields NaN."""
    from oracle import gpzoo_oracle as O
    import gpzoo_b200 as gz
    from gpzoo_b200 import functional as F
    z = {k: torch.from_numpy(v) for k, v in np.load(GOLDEN + "/kernels.npz").items()}
    X, Z = z["X"], z["Z"]
    tol = 1e-12 if dt == torch.float64 else 3e-6
    mk = gz.kernels.batched_Matern32(sigma=1.2, lengthscale=0.8).to(DEV)
    assert relerr(mk(X.to(DEV, dt), Z.to(DEV, dt)), z["matern32"]) < max(tol, 1e-7)          # ctor rounds params to fp32
    assert relerr(mk.forward_distance(O.squared_dist(X, Z).to(DEV, dt)), z["matern32"]) < max(tol, 1e-6)
    # (L,) parameters, ragged sizes, all gradients
    g = torch.Generator().manual_seed(5)
    n1, n2, L = 37, 1030, 3
    x1 = torch.randn(n1, 2, generator=g, dtype=torch.float64)
    x2 = torch.randn(n2, 2, generator=g, dtype=torch.float64)
    sg = 1 + 0.1 * torch.rand(L, generator=g, dtype=torch.float64)
    ls = 0.8 + 0.4 * torch.rand(L, generator=g, dtype=torch.float64)
    Wt = torch.randn(L, n1, n2, generator=g, dtype=torch.float64)
    leaves = [t.clone().requires_grad_(True) for t in (x1, x2, sg, ls)]
    Kref = O.matern32(leaves[0], leaves[1], leaves[2].reshape(L, 1, 1), leaves[3].reshape(L, 1, 1))
    (Kref * Wt).sum().backward()
    dl = [t.detach().to(DEV, dt).requires_grad_(True) for t in (x1, x2, sg, ls)]
    K = F.KernelBuild.apply(dl[0], dl[1], dl[2], dl[3], None, None, None, None, 1.0, 0.0, False, 1)
    tolb = 1e-11 if dt == torch.float64 else 2e-5
    assert relerr(K, Kref) < tolb
    (K * Wt.to(DEV, dt)).sum().backward()
    for i, (r, c) in enumerate(zip(leaves, dl)):
        assert relerr(c.grad, r.grad) < tolb, i
    # coincident points (Kzz with jitter): gradient of the diagonal entries is exactly zero w.r.t. the points
    zz = x1.clone().requires_grad_(True)
    d2 = ((zz[:, None, :] - zz[None, :, :]) ** 2).sum(-1)
    off = ~torch.eye(n1, dtype=torch.bool)
    dsafe = torch.where(off, d2, torch.ones_like(d2)).sqrt() * off                       # sqrt never sees 0
    v = math.sqrt(3.0) * dsafe / ls.reshape(L, 1, 1)
    Kzz_ref = sg.reshape(L, 1, 1) ** 2 * (1 + v) * torch.exp(-v) + 0.05 * torch.eye(n1, dtype=torch.float64)
    Wz = torch.randn(L, n1, n1, generator=g, dtype=torch.float64)
    (Kzz_ref * Wz).sum().backward()
    zd = x1.to(DEV, dt).requires_grad_(True)
    Kzz = F.KernelBuild.apply(zd, zd, sg.to(DEV, dt), ls.to(DEV, dt), None, None, None, None, 1.0, 0.05, False, 1)
    assert relerr(Kzz, Kzz_ref) < tolb
    (Kzz * Wz.to(DEV, dt)).sum().backward()
    assert bool(torch.isfinite(zd.grad).all()) and relerr(zd.grad, zz.grad) < tolb


def test_gemm_split_fp16_route(monkeypatch):
    """functional.gemm sends large fp32 M x M x M products to the split-FP16 kernel (GEMM16_MIN_DIM, 2048 by default); with the
    threshold lowered the same route is checked here at small sizes: transposes, triangular flags, beta = 1 accumulation."""
    from gpzoo_b200 import functional as F
    monkeypatch.setattr(F, "GEMM16_MIN_DIM", 128)
    g = torch.Generator().manual_seed(17)
    b, m = 2, 256
    A = torch.randn(b, m, m, generator=g)
    B = torch.randn(b, m, m, generator=g) * 50.0
    Ad, Bd = A.to(DEV), B.to(DEV)
    for ta in (False, True):
        for tb in (False, True):
            F.clear_step_cache()
            opA = A.double().transpose(1, 2) if ta else A.double()
            opB = B.double().transpose(1, 2) if tb else B.double()
            assert relerr(F.gemm(Ad, Bd, ta=ta, tb=tb), opA @ opB) < 5e-6, (ta, tb)
    F.clear_step_cache()
    Lo, Up = torch.tril(A), torch.triu(B)
    out = F.gemm(Lo.to(DEV), Up.to(DEV), a_tri=1, b_tri=2, d_tri=0)
    assert relerr(out, Lo.double() @ Up.double()) < 5e-6
    C0 = torch.randn(b, m, m, generator=g)
    acc = C0.to(DEV).clone()
    F.gemm(Lo.to(DEV), Lo.to(DEV), tb=True, alpha=-1.0, beta=1.0, out=acc, b_tri=2, d_tri=1)        # lower triangle updated in place
    ref = C0.double() - Lo.double() @ Lo.double().transpose(1, 2)
    assert relerr(torch.tril(acc), torch.tril(ref)) < 5e-6
    assert torch.equal(torch.triu(acc, 1).cpu(), torch.triu(C0, 1))


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_utilities_svgp_forward_and_squared_dist(dt):
    """utilities.svgp_forward (utilities.py:382-397) and _squared_dist (:399-405) on the library's kernels against the oracle's
    restatement, values and gradients."""
    import gpzoo_b200 as gz
    from oracle import gpzoo_oracle as O
    g = torch.Generator().manual_seed(5)
    L, N, M = 3, 70, 20
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    Kxx, W, mu = mk(L, N).abs() + 2, 0.3 * mk(L, N, M), mk(L, M)
    A, B = mk(L, M, M), mk(L, M, M)
    S, Kzz = A @ A.transpose(-1, -2) / M, B @ B.transpose(-1, -2) / M
    ref_in = [t.clone().requires_grad_(True) for t in (Kxx, Kzz, W, mu, S)]
    rm, rc = O.svgp_forward(*ref_in)
    (rm.sum() + (rc ** 2).sum()).backward()
    our_in = [t.to(DEV, dt).requires_grad_(True) for t in (Kxx, Kzz, W, mu, S)]
    m, c = gz.utilities.svgp_forward(*our_in)
    assert m.shape == (L, N, 1) and c.shape == (L, N)
    (m.sum() + (c ** 2).sum()).backward()
    tol = 1e-10 if dt == torch.float64 else 1e-4
    assert relerr(m, rm) < tol and relerr(c, rc) < tol
    for a, b in zip(our_in, ref_in):
        assert relerr(a.grad, b.grad) < tol
    X, Z = mk(40, 2), mk(17, 2)
    Xr, Zr = X.clone().requires_grad_(True), Z.clone().requires_grad_(True)
    wgt = mk(40, 17)
    (O.squared_dist(Xr, Zr) * wgt).sum().backward()
    Xo, Zo = X.to(DEV, dt).requires_grad_(True), Z.to(DEV, dt).requires_grad_(True)
    d2 = gz.utilities._squared_dist(Xo, Zo)
    (d2 * wgt.to(DEV, dt)).sum().backward()
    assert relerr(d2, O.squared_dist(X, Z)) < tol and relerr(Xo.grad, Xr.grad) < tol and relerr(Zo.grad, Zr.grad) < tol
    assert float(gz.utilities._squared_dist(Xo.detach(), Xo.detach()).diagonal().abs().max()) == 0.0


def test_kernel_build_rejects_mismatched_arguments():
    """The C ABI takes raw pointers, so the Python boundary refuses what the kernels would misread: a float32 Z with float64 X,
    group labels outside [0, n_groups), and a Matern kernel under VNNGP (whose fused neighbour kernel evaluates the RBF)."""
    import gpzoo_b200 as gz
    from gpzoo_b200 import _cabi
    X = torch.rand(32, 2, dtype=torch.float64, device=DEV)
    Z = torch.rand(8, 2, dtype=torch.float32, device=DEV)
    k = gz.kernels.NSF_RBF(L=2).to(DEV)
    with pytest.raises(_cabi.GpzError):
        k(X, Z)
    mk = gz.kernels.MGGP_NSF_RBF(L=2, n_groups=3).to(DEV)
    gX = torch.randint(0, 3, (32,), device=DEV)
    gZ = torch.tensor([0, 1, 2, 3, 0, 1, 2, 0], device=DEV)                    # 3 is out of range
    with pytest.raises(IndexError):
        mk(X, Z.double(), gX, gZ)
    out = mk(X, Z.double(), gX.int(), (gZ % 3).int())                           # int32 labels are converted, not reinterpreted
    assert out.shape == (2, 32, 8)
    v = gz.gp.VNNGP(gz.kernels.batched_Matern32(), dim=2, M=8, K=3).to(DEV).double()
    with pytest.raises(NotImplementedError):
        v(X)


@pytest.mark.parametrize("M", [128, 320, 1024, 1600])
def test_fused_chain_matches_generic_fp64(M):
    """csrc/chain.cu (one forward + one backward call, tcgen05 split-TF32 M^3 products) against the Function-per-op chain in
    fp64: every output and, for random incoming gradients of every output, every input gradient.  gKzz is compared after
    symmetrisation (the fused chain returns the un-symmetrised representative)."""
    from gpzoo_b200 import functional as F
    g = torch.Generator().manual_seed(M)
    L = 3
    Q = torch.randn(L, M, M, generator=g, dtype=torch.float64)
    K = (Q @ Q.transpose(1, 2) / M + torch.eye(M, dtype=torch.float64)).to(DEV)
    raw = (0.3 * torch.randn(L, M, M, generator=g, dtype=torch.float64)).to(DEV)
    mu = torch.randn(L, M, generator=g, dtype=torch.float64).to(DEV)
    w = [torch.randn(L, M, M, generator=g, dtype=torch.float64).tril().to(DEV) for _ in range(4)]
    wq, wk = torch.randn(L, M, generator=g, dtype=torch.float64).to(DEV), torch.randn(L, generator=g, dtype=torch.float64).to(DEV)

    def loss(Lc, Linv, Lu, T, q, kl, dt):
        c = lambda t: t.to(dt)
        return ((Lc * c(w[0])).sum() + (Linv * c(w[1])).sum() + (Lu * c(w[2])).sum() + (T * c(w[3])).sum() + (q * c(wq)).sum()
                + (kl * c(wk)).sum())
    # generic fp64
    K64, r64, m64 = (t.clone().requires_grad_(True) for t in (K, raw, mu))
    Lc, Linv = F.CholeskyInverse.apply(K64)
    Lu = F.LowerCholesky.apply(r64)
    T, q = F.Whiten.apply(Linv, Lu, m64)
    kl = F.MvnKL.apply(T, q, Lc, Lu)
    ref = (Lc, Linv, Lu, T, q, kl)
    loss(*ref, torch.float64).backward()
    # fused fp32
    assert F.chain_ok(torch.float32, M)
    K32, r32, m32 = (t.float().requires_grad_(True) for t in (K, raw, mu))
    out = F.SvgpChain.apply(K32, r32, m32, False)
    loss(*out, torch.float32).backward()
    for name, a, b in zip(("Lc", "Linv", "Lu", "T", "q", "kl"), out, ref):
        assert relerr(a, b) < 2e-5, (name, relerr(a, b))
    sym = lambda t: 0.5 * (t + t.transpose(1, 2))
    assert relerr(sym(K32.grad), sym(K64.grad)) < 1e-4, relerr(sym(K32.grad), sym(K64.grad))
    assert relerr(r32.grad, r64.grad) < 1e-4 and relerr(m32.grad, m64.grad) < 1e-4
    assert torch.equal(K32.detach(), K.float())                                 # consume=False: the input survives
    assert torch.equal(r32.grad.triu(1), torch.zeros_like(r32.grad))
    # partial incoming gradients (None for the rest) and a second backward over the same graph
    K32b = K.float().requires_grad_(True)
    o2 = F.SvgpChain.apply(K32b, raw.float(), mu.float(), False)
    o2[5].sum().backward(retain_graph=True)
    g1 = K32b.grad.clone()
    K32b.grad = None
    o2[5].sum().backward()
    assert relerr(K32b.grad, g1) < 1e-6
    K64b = K.clone().requires_grad_(True)
    Lcb, Linvb = F.CholeskyInverse.apply(K64b)
    Tb, qb = F.Whiten.apply(Linvb, F.LowerCholesky.apply(raw), mu)
    F.MvnKL.apply(Tb, qb, Lcb, F.LowerCholesky.apply(raw)).sum().backward()
    assert relerr(sym(g1), sym(K64b.grad)) < 1e-4
