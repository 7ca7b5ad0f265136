"""fp64 stage-by-stage check at config-2 conditioning (M = 1024, L = 10): where does the 1e-9 come from?"""
import sys, torch
sys.path.insert(0, '.')
import gpzoo_b200 as gz
from gpzoo_b200 import functional as F, synthetic
from oracle import gpzoo_oracle as O
dev = 'cuda'
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
prob = synthetic.nsf_problem(N=1024, M=1024, L=10, G=50, E=1, seed=1, coord_scale=100.0, lengthscale=1.7, jitter=1e-1)
prob["Lu_raw"] = 0.05 * synthetic.hash_uniform(10, 1024, 1024, salt=7)
Z, X, sg, ls = (prob[k].to(dev) for k in ("Z", "X", "sigma", "lengthscale"))
kern = gz.kernels.NSF_RBF(L=10); kern.sigma, kern.lengthscale = torch.nn.Parameter(sg), torch.nn.Parameter(ls)
with torch.no_grad():
    Kzz = kern(Z, Z, _jitter=0.1); Kzx = kern(Z, X)
    Kzz_o = O.with_jitter(O.nsf_rbf(prob["Z"], prob["Z"], prob["sigma"], prob["lengthscale"]), 0.1)
    Kzx_o = O.nsf_rbf(prob["Z"], prob["X"], prob["sigma"], prob["lengthscale"])
    print("Kzz", rel(Kzz, Kzz_o), "Kzx", rel(Kzx, Kzx_o))
    Lc, Linv = F.CholeskyInverse.apply(Kzz.clone())
    Lr = torch.linalg.cholesky(Kzz_o)
    Lir = torch.linalg.inv(Lr)
    print("Lc", rel(Lc, Lr), "Linv", rel(Linv, Lir), "per factor Linv", [("%.0e" % rel(Linv[l], Lir[l])) for l in range(10)])
    Lu_o = O.lower_cholesky_transform(prob["Lu_raw"])
    Lu = F.LowerCholesky.apply(prob["Lu_raw"].to(dev))
    print("Lu", rel(Lu, Lu_o))
    T, q = F.Whiten.apply(Linv, Lu, prob["mu"].to(dev))
    print("T", rel(T, Lir @ Lu_o), "q", rel(q, (Lir @ prob["mu"].unsqueeze(-1)).squeeze(-1)))
    A = F.gemm(Linv, Kzx, a_tri=1)
    print("A", rel(A, torch.linalg.solve_triangular(Lr, Kzx_o, upper=False)))
    kl = F.MvnKL.apply(T, q, Lc, Lu)
    print("kl", rel(kl, O.mvn_kl(prob["mu"], Lu_o, Lr)))
    Kxx = kern(X, X, diag=True)
    mean, var = F.Predict.apply(Kxx, Kzx, Linv, T, q, None)
    mo, vo, _, _ = O.svgp(O.rbf_diag(prob["X"], prob["sigma"]), Kzx_o, O.nsf_rbf(prob["Z"], prob["Z"], prob["sigma"], prob["lengthscale"]), prob["mu"], prob["Lu_raw"], 0.1, 1e-6)
    print("mean", rel(mean, mo), "var", rel(var, vo))
