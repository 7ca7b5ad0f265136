"""Does CUDA-graph replay of the whole ELBO fwd+bwd step (config 2) beat eager launches?  (probe, one GPU)"""
import sys, torch
sys.path.insert(0, '.')
import bench
from gpzoo_b200 import functional, synthetic
functional.set_sync_checks(False)
dev, dt = torch.device('cuda', 0), torch.float32
c = bench.CFG
prob = synthetic.nsf_problem(N=c["N"], M=c["M"], L=c["L"], G=c["G"], E=c["E"], seed=c["seed"], coord_scale=c["coord_scale"],
                             lengthscale=c["lengthscale"], jitter=c["jitter"], dtype=dt)
model, shared = bench.build_model(prob, dt, dev)
X, y, eps = prob["X"].to(dev), prob["y"].to(dev), prob["eps"].to(dev)

def step():
    for p in model.parameters():
        p.grad = None
    elbo = model.elbo(X, y, E=c["E"], eps=eps)
    (-elbo).backward()
    return elbo.detach()

def timeit(fn, n=20):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

for _ in range(3): step()
print("eager ms/step", timeit(step))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
g.replay(); torch.cuda.synchronize()
ref = step()
g.replay(); torch.cuda.synchronize()
print("graph elbo", float(out), "eager elbo", float(ref))
print("graph ms/step", timeit(g.replay))
