"""Single-process emulation of the 2-rank data-parallel step: two half shards with kl_weight = 1/2, gradients summed, against the full step."""
import sys, torch
sys.path.insert(0, '.')
import bench
from gpzoo_b200 import functional
functional.set_sync_checks(False)
dev, dt = torch.device('cuda'), torch.float32
c = dict(bench.CONFIGS[2]); c["N"] = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prob = bench.make_problem(c, c["N"], c["seed"], dt)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))
names = ["Z", "sigma", "lengthscale", "mu", "Lu", "W"]
def run(sl, w):
    p = dict(prob); p["V"] = prob["V"][sl].contiguous()
    m, sh = bench.build_model(c, p, dt, dev)
    e = m.elbo(prob["X"][sl].to(dev), prob["y"][:, sl].contiguous().to(dev), E=1, eps=prob["eps"][:, :, sl].contiguous().to(dev), kl_weight=w)
    (-e).backward()
    return e.detach(), [q.grad.clone() for q in sh]
for trial in range(2):
    e_full, g_full = run(slice(0, c["N"]), 1.0)
    h = c["N"] // 2
    e0, g0 = run(slice(0, h), 0.5)
    e1, g1 = run(slice(h, c["N"]), 0.5)
    print("elbo", rel(e0 + e1, e_full), {n: "%.1e" % rel(a + b, f) for n, a, b, f in zip(names, g0, g1, g_full)})
    e_full2, g_full2 = run(slice(0, c["N"]), 1.0)
    print("full again", {n: "%.1e" % rel(a, f) for n, a, f in zip(names, g_full2, g_full)})
