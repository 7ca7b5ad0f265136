"""The Poisson likelihood kernel alone at config 2's shape (ncu target)."""
import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
torch.manual_seed(0)
G, Fn, B, E = 2000, 10, 32768, 1
dev = 'cuda'
y = torch.poisson(torch.full((G, B), 0.3, device=dev))
W = torch.rand(G, Fn, device=dev); V = torch.ones(B, device=dev)
mean = 0.3 * torch.randn(Fn, B, device=dev); var = 0.1 + torch.rand(Fn, B, device=dev); eps = torch.randn(E, Fn, B, device=dev)
for _ in range(4):
    ll = F.PoissonLL.apply(y, None, W, V, mean, var, eps, Fn, 1e-6, True, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ll = F.PoissonLL.apply(y, None, W, V, mean, var, eps, Fn, 1e-6, True, True)
e1.record(); torch.cuda.synchronize()
print("poisson ms", e0.elapsed_time(e1) / 20, float(ll))
