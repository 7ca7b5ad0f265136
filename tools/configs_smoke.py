"""Time one fwd+bwd ELBO step at the shapes of BASELINE.json configs 1, 3, 4, 5 (per-GPU minibatch) on one B200."""
import sys, time, torch
sys.path.insert(0, '.')
import gpzoo_b200 as gz
from gpzoo_b200 import synthetic, functional
functional.set_sync_checks(False)
dev, dt = 'cuda', torch.float32
P = lambda t: torch.nn.Parameter(t.to(dev, dt) if t.is_floating_point() else t.to(dev))

def timeit(fn, n=5, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

def nsf_model(prob, gp_cls, kern, **gpkw):
    L, M = prob["mu"].shape
    kern.sigma, kern.lengthscale = P(prob["sigma"]), P(prob["lengthscale"])
    gp = gp_cls(kern, dim=2, M=M, jitter=prob["jitter"], **gpkw)
    gp.Z, gp.mu, gp.Lu = P(prob["Z"]), P(prob["mu"]), P(prob["Lu_raw"])
    return gp

def step(model, *a, **k):
    for p in model.parameters(): p.grad = None
    e = model.elbo(*a, **k); (-e).backward(); return e

# config 1: SVGP regression N=2000 M=100 E=20
pr = synthetic.regression_problem(N=2000, M=100, E=20, seed=0, dtype=dt, device=dev)
k = gz.kernels.RBF(); k.sigma, k.lengthscale = P(pr["sigma"]), P(pr["lengthscale"])
gp = gz.gp.SVGP(k, dim=2, M=100, jitter=pr["jitter"]); gp.Z, gp.mu, gp.Lu = P(pr["Z"]), P(pr["mu"]), P(pr["Lu_raw"])
m1 = gz.likelihoods.GaussianLikelihood(gp).to(dev)
print("config1 SVGP-Gaussian N=2000 M=100 E=20: %.3f ms/step" % timeit(lambda: step(m1, pr["X"], pr["y"], E=20, eps=pr["eps"])), flush=True)

# config 3: VNNGP K=8 N=4000 M=1000 L=10 G=2000 E=10
pr = synthetic.nsf_problem(N=4000, M=1000, L=10, G=2000, E=10, seed=3, coord_scale=2.0, lengthscale=0.25, jitter=1e-2, dtype=dt, device=dev)
gp = nsf_model(pr, gz.gp.VNNGP, gz.kernels.NSF_RBF(L=10), K=8)
m3 = gz.likelihoods.NSF2(gp, pr["y"][:, :1], L=10); m3.W, m3.V = P(pr["W"]), P(pr["V"])
print("config3 NSF-VNNGP K=8 N=4000 M=1000 L=10 G=2000 E=10: %.3f ms/step" % timeit(lambda: step(m3, pr["X"], pr["y"], E=10, eps=pr["eps"])), flush=True)
del m3, gp, pr; torch.cuda.empty_cache()

# config 4: MGGP-SVGP 10 groups, M=2048, per-GPU minibatch 8192 of N=100k (here: one rank's share)
pr = synthetic.nsf_problem(N=8192, M=2048, L=10, G=2000, E=1, seed=4, coord_scale=100.0, lengthscale=3.0, jitter=1e-1, dtype=dt, device=dev, n_groups=10)
kern = gz.kernels.MGGP_NSF_RBF(L=10, n_groups=10); kern.set_group_distances(pr["group_distances"].float().cpu())
kern.embedding = torch.nn.Parameter(kern.embedding.to(dev, dt), requires_grad=False); kern.group_diff_param = P(pr["gdp"])
gp = nsf_model(pr, gz.gp.MGGP_SVGP, kern, n_groups=10); gp.groupsZ = torch.nn.Parameter(pr["groupsZ"].to(dev), requires_grad=False)
m4 = gz.likelihoods.NSF2(gp, pr["y"][:, :1], L=10); m4.W, m4.V = P(pr["W"]), P(pr["V"])
print("config4 NSF-MGGP_SVGP ng=10 M=2048 B=8192/GPU: %.3f ms/step" % timeit(lambda: step(m4, pr["X"], pr["y"], E=1, eps=pr["eps"], groupsX=pr["groupsX"])), flush=True)
del m4, gp, pr; torch.cuda.empty_cache()

# config 5: Hybrid_NSF2 M=4096, T=10, per-GPU minibatch 16384
pr = synthetic.nsf_problem(N=16384, M=4096, L=10, G=2000, E=1, seed=5, coord_scale=100.0, lengthscale=1.7, jitter=1e-1, dtype=dt, device=dev)
gp = nsf_model(pr, gz.gp.SVGP, gz.kernels.NSF_RBF(L=10))
prior = gz.gp.GaussianPrior(pr["y"], L=10).to(dev)
m5 = gz.likelihoods.Hybrid_NSF2(gp, prior, pr["y"][:, :1], L=10, T=10).to(dev); m5.V = P(pr["V"])
print("config5 Hybrid_NSF2 M=4096 T=10 B=16384/GPU: %.3f ms/step" % timeit(lambda: step(m5, pr["X"], pr["y"], E=1), n=3, w=1), flush=True)
print("peak mem GB", torch.cuda.max_memory_allocated() / 1e9)
functional.check_cholesky_info()
