"""Small ragged shapes of the Poisson likelihood call in one process (compute-sanitizer target): odd row lengths, gathered columns,
integer counts, several samples, G below one 16-gene block, F = 16, a spot count that is not a multiple of anything."""
import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
torch.manual_seed(0)
dev = 'cuda'
for (G, Fn, Ntot, B, E, ydt) in [(45, 5, 300, 300, 3, None), (7, 16, 333, 333, 1, torch.uint8), (130, 10, 401, 257, 2, torch.int16),
                                 (33, 3, 1000, 999, 1, torch.int32), (2000, 10, 515, 515, 1, None), (17, 12, 64, 64, 2, torch.uint8)]:
    y = torch.poisson(torch.full((G, Ntot), 0.4, device=dev))
    y = y if ydt is None else y.to(ydt)
    idx = torch.randperm(Ntot, device=dev)[:B] if B != Ntot else None
    W = torch.rand(G, Fn, device=dev, requires_grad=True); V = torch.ones(Ntot, device=dev, requires_grad=True)
    mean = (0.3 * torch.randn(Fn, B, device=dev)).requires_grad_(True); var = (0.1 + torch.rand(Fn, B, device=dev)).requires_grad_(True)
    eps = torch.randn(E, Fn, B, device=dev)
    ll = F.PoissonLL.apply(y, idx, W, V, mean, var, eps, Fn // 2, 5e-2, True, True)
    ll.backward()
    torch.cuda.synchronize()
    print(G, Fn, Ntot, B, E, ydt, float(ll), bool(torch.isfinite(W.grad).all() and torch.isfinite(mean.grad).all()))
print("done")
