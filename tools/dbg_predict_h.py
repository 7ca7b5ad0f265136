"""Per-tensor error of the fp32 tensor-core predict paths (fp16x3 / tf32x3) and the CUDA-core fp32 path vs fp64."""
import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
from tests.helpers import relerr
DEV = 'cuda'
g = torch.Generator().manual_seed(13)
L, M, N = 2, 192, 768
Z = torch.rand(M, 2, generator=g, dtype=torch.float64) * 10
X = torch.rand(N, 2, generator=g, dtype=torch.float64) * 10
sg = torch.tensor([1.0, 2.5], dtype=torch.float64)
ls = torch.tensor([0.7, 1.0], dtype=torch.float64)
mu = torch.randn(L, M, generator=g, dtype=torch.float64)
Lur = 0.1 * torch.randn(L, M, M, generator=g, dtype=torch.float64)
wm = torch.randn(L, N, generator=g, dtype=torch.float64)
wv = torch.randn(L, N, generator=g, dtype=torch.float64)
res = {}
for name, dt, arith in (("f64", torch.float64, None), ("fp16x3", torch.float32, "fp16x3"), ("tf32x3", torch.float32, "tf32x3"),
                        ("simt32", torch.float32, "simt")):
    d = lambda t: t.to(DEV, dt)
    Zd, mud, Lud = d(Z).requires_grad_(True), d(mu).requires_grad_(True), d(Lur).requires_grad_(True)
    lsd, sgd = d(ls).requires_grad_(True), d(sg).requires_grad_(True)
    F.USE_TENSOR_CORES = arith != "simt"
    Kzz = F.KernelBuild.apply(Zd, Zd, sgd, lsd, None, None, None, None, 1.0, 0.05)
    Lc, Linv = F.CholeskyInverse.apply(Kzz)
    Lu = F.LowerCholesky.apply(Lud)
    T, q = F.Whiten.apply(Linv, Lu, mud)
    Kxx = (sgd ** 2)[:, None].expand(-1, N).contiguous()
    hook = {}
    if arith == "fp16x3":
        Kzx, Kh, Kl, sK = F.KernelBuildH.apply(Zd, d(X), sgd, lsd, None, None, None, None, 1.0, 0.0)
        mean, var = F.PredictH.apply(Kxx, Kzx, Linv, T, q, Kh, Kl, sK)
    else:
        out = F.KernelBuild.apply(Zd, d(X), sgd, lsd, None, None, None, None, 1.0, 0.0, arith == "tf32x3")
        Kzx, Kzx_lo = out if isinstance(out, tuple) else (out, None)
        mean, var = F.Predict.apply(Kxx, Kzx, Linv, T, q, Kzx_lo)
    Kzx.register_hook(lambda g_: hook.__setitem__("gKzx", g_.clone()))
    Linv.register_hook(lambda g_: hook.__setitem__("gLinv", g_.clone()))
    T.register_hook(lambda g_: hook.__setitem__("gT", g_.clone()))
    q.register_hook(lambda g_: hook.__setitem__("gq", g_.clone()))
    ((mean * d(wm)).sum() + (var * d(wv)).sum()).backward()
    res[name] = dict(mean=mean, var=var, gZ=Zd.grad, gmu=mud.grad, gLu=Lud.grad, gls=lsd.grad, gsg=sgd.grad, **hook)
for name in ("fp16x3", "tf32x3", "simt32"):
    print(name, {k: f"{relerr(res[name][k], res['f64'][k]):.2e}" for k in res["f64"]})
