"""Per-C-ABI-call time of one fwd+bwd step at the per-GPU shapes of BASELINE.json configs 4 and 5 (CUDA events per call)."""
import sys, torch
sys.path.insert(0, '.')
import gpzoo_b200 as gz
from gpzoo_b200 import synthetic, functional, _cabi
functional.set_sync_checks(False)
dev, dt = 'cuda', torch.float32
P = lambda t: torch.nn.Parameter(t.to(dev, dt) if t.is_floating_point() else t.to(dev))

def run(name, N, M, steps=3):
    pr = synthetic.nsf_problem(N=N, M=M, L=10, G=2000, E=1, seed=5, coord_scale=100.0, lengthscale=1.7 if M == 4096 else 3.0, jitter=1e-1, dtype=dt, device=dev)
    kern = gz.kernels.NSF_RBF(L=10); kern.sigma, kern.lengthscale = P(pr["sigma"]), P(pr["lengthscale"])
    gp = gz.gp.SVGP(kern, dim=2, M=M, jitter=pr["jitter"]); gp.Z, gp.mu, gp.Lu = P(pr["Z"]), P(pr["mu"]), P(pr["Lu_raw"])
    m = gz.likelihoods.NSF2(gp, pr["y"][:, :1], L=10); m.W, m.V = P(pr["W"]), P(pr["V"])
    def step():
        for p in m.parameters(): p.grad = None
        e = m.elbo(pr["X"], pr["y"], E=1, eps=pr["eps"]); (-e).backward()
    for _ in range(2): step()
    torch.cuda.synchronize()
    _cabi.profile = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): step()
    e1.record(); torch.cuda.synchronize()
    prof, _cabi.profile = _cabi.profile, None
    tot = e0.elapsed_time(e1) / steps
    per = {k: sum(a.elapsed_time(b) for a, b in v) / steps for k, v in prof.items()}
    print(f"{name}: {tot:.2f} ms/step;", {k: round(v, 2) for k, v in sorted(per.items(), key=lambda kv: -kv[1]) if v > 0.05 * tot / 10})

run("config4-like SVGP M=2048 B=8192", 8192, 2048)
run("config5-like SVGP M=4096 B=16384", 16384, 4096)
