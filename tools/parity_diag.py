"""Where does the fp32 error of the tensor-core path come from?  For a few problem shapes: run the step in fp64 on the GPU
(truth; the fp64 path is pinned to the reference goldens at 1e-10) and in fp32 with (a) CUDA-core GEMMs, (b) split-TF32,
(c) split-FP16 tcgen05 GEMMs, and print the relative L2 error of the ELBO pieces and of every gradient.  For the small shapes
the reference's own fp32 floor (oracle port on the CPU, stock torch.cdist and exact cdist) is printed beside it.
Dev tool (GPU box): python tools/parity_diag.py [case ...]"""
import sys

import torch

sys.path.insert(0, '.')
import gpzoo_b200 as gz  # noqa: E402
from gpzoo_b200 import functional, synthetic  # noqa: E402

dev = 'cuda'
CASES = {
    "smoke": dict(N=512, M=64, L=3, G=24, E=2, seed=1, coord_scale=2.0, jitter=1e-2),
    "tc128": dict(N=1024, M=128, L=3, G=32, E=1, seed=2, coord_scale=2.0, jitter=1e-2),
    "mid256": dict(N=4096, M=256, L=4, G=200, E=1, seed=3, coord_scale=100.0, lengthscale=1.7, jitter=1e-1),
    "cfg2cond": dict(N=1024, M=1024, L=10, G=2000, E=1, seed=1, coord_scale=100.0, lengthscale=1.7, jitter=1e-1),
    "cfg2cond8k": dict(N=8192, M=1024, L=10, G=2000, E=1, seed=1, coord_scale=100.0, lengthscale=1.7, jitter=1e-1),
}


def rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def run_gpu(prob, dt):
    P = lambda t: torch.nn.Parameter(t.to(dev, dt))
    L, M = prob["mu"].shape
    kern = gz.kernels.NSF_RBF(L=L)
    kern.sigma, kern.lengthscale = P(prob["sigma"]), P(prob["lengthscale"])
    gp = gz.gp.SVGP(kern, dim=prob["X"].shape[1], M=M, jitter=prob["jitter"])
    gp.Z, gp.mu, gp.Lu = P(prob["Z"]), P(prob["mu"]), P(prob["Lu_raw"])
    model = gz.likelihoods.NSF2(gp, prob["y"][:, :1], L=L)
    model.W, model.V = P(prob["W"]), P(prob["V"])
    elbo, parts = model.elbo(prob["X"].to(dev, dt), prob["y"].to(dev, dt), E=prob["eps"].shape[0], eps=prob["eps"].to(dev, dt),
                             return_parts=True)
    elbo.backward()
    out = dict(elbo=elbo, ll=parts["ll"], kl=parts["kl"], mean=parts["mean"], var=parts["var"])
    out.update({"d" + k: v.grad for k, v in dict(Z=gp.Z, sigma=kern.sigma, ls=kern.lengthscale, mu=gp.mu, Lu=gp.Lu, W=model.W,
                                                 V=model.V).items()})
    return {k: v.detach().double().cpu() for k, v in out.items()}


def run_oracle(prob, dt, exact):
    from oracle import gpzoo_oracle as O
    O.set_exact_cdist(exact)
    p = O.NSFParams(**{k: prob[k].to(dt).clone() for k in ("Z", "sigma", "lengthscale", "mu", "Lu_raw", "W", "V")}, jitter=prob["jitter"])
    ref, g = O.value_and_grads(lambda: O.nsf_svgp_terms(p, prob["X"].to(dt), prob["y"].to(dt), prob["eps"].to(dt)), p.leaves())
    O.set_exact_cdist(False)
    out = dict(elbo=ref["elbo"], ll=ref["ll"], kl=ref["kl"], mean=ref["mean"], var=ref["var"])
    out.update(dZ=g["Z"], dsigma=g["sigma"], dls=g["lengthscale"], dmu=g["mu"], dLu=g["Lu_raw"], dW=g["W"], dV=g["V"])
    return {k: v.detach().double() for k, v in out.items()}


def chol_accuracy(name):
    """Lc and Linv of the jittered Kzz in fp32 (ours vs torch.linalg on the same GPU) against fp64."""
    prob = synthetic.nsf_problem(**CASES[name])
    from oracle import gpzoo_oracle as O
    Kzz = O.with_jitter(O.nsf_rbf(prob["Z"], prob["Z"], prob["sigma"], prob["lengthscale"]), prob["jitter"]).to(dev)
    Lc64 = torch.linalg.cholesky(Kzz)
    eye = torch.eye(Kzz.shape[-1], dtype=torch.float64, device=dev).expand_as(Kzz)
    Li64 = torch.linalg.solve_triangular(Lc64, eye, upper=False)
    K32 = Kzz.float()
    Lc, Linv = functional.CholeskyInverse.apply(K32.clone())
    Lt = torch.linalg.cholesky(K32)
    Lit = torch.linalg.solve_triangular(Lt, eye.float(), upper=False)
    print(f"[chol {name}] cond={float(torch.linalg.cond(Kzz[0])):.0f}  ours: Lc {rel(Lc, Lc64):.1e} Linv {rel(Linv, Li64):.1e} | torch fp32: Lc {rel(Lt, Lc64):.1e} "
          f"Linv {rel(Lit, Li64):.1e} | ours Linv*Lc-I {float((Linv.double() @ Lc64 - eye).norm() / eye.norm()):.1e} torch {float((Lit.double() @ Lc64 - eye).norm() / eye.norm()):.1e}")


def main():
    names = sys.argv[1:] or ["smoke", "tc128", "mid256", "cfg2cond"]
    for name in [n[5:] for n in names if n.startswith("chol:")]:
        chol_accuracy(name)
    names = [n for n in names if not n.startswith("chol:")]
    for name in names:
        prob = synthetic.nsf_problem(**CASES[name])
        functional.USE_TENSOR_CORES = True
        functional.TENSOR_CORE_ARITH = "fp16x3"
        truth = run_gpu(prob, torch.float64)
        rows = {}
        for label, tc, arith in (("simt", False, "fp16x3"), ("tf32x3", True, "tf32x3"), ("fp16x3", True, "fp16x3")):
            functional.USE_TENSOR_CORES, functional.TENSOR_CORE_ARITH = tc, arith
            rows[label] = run_gpu(prob, torch.float32)
        functional.USE_TENSOR_CORES, functional.TENSOR_CORE_ARITH = True, "fp16x3"
        if CASES[name]["M"] <= 256:
            o64 = run_oracle(prob, torch.float64, False)
            print(f"[{name}] gpu fp64 vs oracle fp64:", {k: f"{rel(truth[k], o64[k]):.1e}" for k in truth})
            rows["ref32"] = run_oracle(prob, torch.float32, False)
            rows["ref32x"] = run_oracle(prob, torch.float32, True)
        keys = list(truth)
        print(f"[{name}] {CASES[name]}")
        print("   %-8s" % "" + " ".join("%8s" % k for k in keys))
        for label, r in rows.items():
            print("   %-8s" % label + " ".join("%8.1e" % rel(r[k], truth[k]) for k in keys))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
