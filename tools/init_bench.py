"""Initialisation pipeline at config 2's size: regularized_nmf (Kullback-Leibler multiplicative updates, the notebooks' call) on the
device against the reference's sklearn call on the host cores, same start (init='random'), fixed iteration count (tol=0)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gpzoo_b200 import initialisation as I

N, G, L = 32768, 2000, 10
it_gpu, it_cpu = 100, 5
rng = np.random.RandomState(0)
Y = rng.poisson(0.3, size=(N, G)).astype(np.float32)
kw = dict(solver="mu", beta_loss="kullback-leibler", init="random", random_state=0, tol=0.0)
Yd = torch.from_numpy(Y).cuda()
I.nmf(Yd, L, max_iter=2, **kw); torch.cuda.synchronize()
t0 = time.perf_counter(); W, H, n = I.nmf(Yd, L, max_iter=it_gpu, return_n_iter=True, **kw); torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / it_gpu
from sklearn.decomposition import NMF
import warnings; warnings.filterwarnings("ignore")
t0 = time.perf_counter(); m = NMF(L, max_iter=it_cpu, **kw); Wc = m.fit_transform(Y); t_cpu = (time.perf_counter() - t0) / it_cpu
W5, H5 = I.nmf(Yd, L, max_iter=it_cpu, **kw)
rel = float((W5.cpu() - torch.from_numpy(Wc)).norm() / torch.from_numpy(Wc).norm())
t0 = time.perf_counter(); Z, inertia = I.kmeans_inducing(torch.rand(N, 2, device='cuda') * 200 - 100, 1024, n_iter=20); torch.cuda.synchronize()
t_km = time.perf_counter() - t0
print(f"nmf KL-MU N={N} G={G} L={L} fp32: device {t_gpu * 1e3:.1f} ms/iteration, sklearn ({torch.get_num_threads()} host threads) "
      f"{t_cpu * 1e3:.0f} ms/iteration -> {t_cpu / t_gpu:.0f}x; W after {it_cpu} iterations rel diff {rel:.1e}; "
      f"k-means M=1024 (seeding + <=20 Lloyd iterations) {t_km:.2f} s")
