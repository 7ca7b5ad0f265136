"""Cholesky + inverse timing at the benchmark size: cluster kernel (with its phase clocks) vs the tensor-core recursion,
for L = 10 factors and for the 1-2 factors a rank would own if the chain were sharded by factor."""
import ctypes, sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F, _cabi
F.set_sync_checks(False)
torch.manual_seed(0)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for M in [int(a) for a in sys.argv[1:]] or [1024]:
    for L in (10, 2, 1):
        Q = torch.randn(L, M, M, device='cuda')
        K = Q @ Q.transpose(1, 2) / M + torch.eye(M, device='cuda')
        ref = torch.linalg.cholesky(K.double())
        out = {}
        for name, thr in (("cluster", 1 << 30), ("tc", 0)):
            F.CHOL_TC_MIN_M = thr
            if name == "cluster" and M > 1536:
                continue
            t = timeit(lambda: F.CholeskyInverse.apply(K))
            Lc, Linv = F.CholeskyInverse.apply(K)
            err = float((Lc.double() - ref).norm() / ref.norm())
            erri = float((Linv.double() @ ref - torch.eye(M, device='cuda', dtype=torch.float64)).norm() / M ** 0.5)
            out[name] = (t, err, erri)
        tt = timeit(lambda: torch.linalg.cholesky(K))
        line = f"M={M} L={L}: " + "  ".join(f"{k} {v[0]:.3f} ms (Lc err {v[1]:.1e}, |Linv Lc - I| {v[2]:.1e})" for k, v in out.items())
        print(line + f"  torch.linalg.cholesky alone {tt:.3f} ms")
    if M <= 1536:
        F.CHOL_TC_MIN_M = 1 << 30
        dbg = torch.zeros(10, dtype=torch.int64, device='cuda')
        _cabi.lib().gpz_chol_debug_(ctypes.c_void_p(dbg.data_ptr()))
        F.CholeskyInverse.apply(K)
        torch.cuda.synchronize()
        _cabi.lib().gpz_chol_debug_(ctypes.c_void_p(0))
        c = dbg.cpu().tolist()
        print(f"  cluster phases (clock64 of CTA 0, L={L}): leaf {c[0]} panel {c[1]} trail+leaf {c[2]} inverse {c[3]}  total {sum(c[:4])}; first leaf: load {c[4]} diag0 {c[5]} products {c[6]} diag1 {c[7]} X21 {c[8]} store {c[9]}")

# does the time depend on the VALUES?  config 2's Kzz is 1.1 I plus mostly tiny (many denormal) entries
from gpzoo_b200 import synthetic
import gpzoo_b200 as gz
pr = synthetic.nsf_problem(N=64, M=1024, L=10, G=4, E=1, seed=1, coord_scale=100.0, lengthscale=1.7, jitter=1e-1, dtype=torch.float32, device='cuda')
kern = gz.kernels.NSF_RBF(L=10)
kern.sigma, kern.lengthscale = torch.nn.Parameter(pr["sigma"]), torch.nn.Parameter(pr["lengthscale"])
with torch.no_grad():
    Kzz = kern(pr["Z"], pr["Z"], _jitter=0.1)
F.CHOL_TC_MIN_M = 1 << 30
den = float(((Kzz != 0) & (Kzz.abs() < 1.2e-38)).float().mean())
print(f"config-2 Kzz: cluster {timeit(lambda: F.CholeskyInverse.apply(Kzz)):.3f} ms   (fraction of denormal entries {den:.3f}, zeros {float((Kzz == 0).float().mean()):.3f})")
Kz2 = torch.where(Kzz.abs() < 1e-30, torch.zeros_like(Kzz), Kzz)
print(f"config-2 Kzz, tiny entries flushed to 0: cluster {timeit(lambda: F.CholeskyInverse.apply(Kz2)):.3f} ms")
