"""Split-FP16 tcgen05 GEMM vs fp64 on small / ragged shapes (both operand layouts), plus timing at config-2 shapes."""
import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
torch.manual_seed(0)
dev = 'cuda'
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
for bk in (1, 0):
    for (bsz, m, n, k) in ((1, 128, 256, 32), (1, 128, 256, 64), (2, 256, 512, 128), (2, 200, 328, 136), (1, 384, 1024, 512)):
        A = torch.randn(bsz, m, k, device=dev) * 3.0
        B = torch.randn((bsz, n, k) if bk else (bsz, k, n), device=dev) * 0.01
        ref = A.double() @ (B.double().transpose(1, 2) if bk else B.double())
        Ap, Bp = F.split16(A), F.split16(B)
        for nt in (1, 3):
            out = F.umma_gemm16(Ap, Bp, bk, n_terms=nt)
            torch.cuda.synchronize()
            print(f"bk={bk} b{bsz} m{m} n{n} k{k} terms={nt}: rel={rel(out, ref):.3e}")
        sd = torch.full((bsz,), 2.0 ** 6, device=dev)
        (Dh, Dl), amax = F.umma_gemm16(Ap, Bp, bk, out_planes=True, out_scale=sd, want_amax=True)
        rec = (Dh.double() + Dl.double()) / sd.double()[:, None, None]
        print(f"   planes rel={rel(rec, ref):.3e}  amax={amax.tolist()} ref_amax={ref.abs().amax((1, 2)).tolist()}")
# triangular A, split-K
bsz, m, n, k = 2, 512, 768, 512
A = torch.tril(torch.randn(bsz, m, k, device=dev)); B = torch.randn(bsz, k, n, device=dev)
out = F.umma_gemm16(F.split16(A), F.split16(B), 0, a_tri=1)
print("a_tri=1 rel", rel(out, A.double() @ B.double()))
Au = torch.triu(torch.randn(bsz, m, k, device=dev))
out = F.umma_gemm16(F.split16(Au), F.split16(B), 0, a_tri=2)
print("a_tri=2 rel", rel(out, Au.double() @ B.double()))
X = torch.randn(bsz, 256, 4096, device=dev); Y = torch.randn(bsz, 256, 4096, device=dev)
out = F.umma_gemm16(F.split16(X), F.split16(Y), 1, d_tri=1, splitk=4)
print("NT tril splitk rel", rel(out, torch.tril(X.double() @ Y.double().transpose(1, 2))))
# transposed planes
h, l, hT, lT, s = F.split16(A, transpose=True)
print("transpose planes ok", bool(torch.equal(hT, h.transpose(1, 2).contiguous()) and torch.equal(lT, l.transpose(1, 2).contiguous())))

if len(sys.argv) > 1 and sys.argv[1] == "time":
    L, M, N = 10, 1024, 32768
    A = torch.tril(torch.randn(L, M, M, device=dev)); B = torch.randn(L, M, N, device=dev)
    Ap, Bp = F.split16(A), F.split16(B)
    Alo, Blo = F.tf32_lo(A), F.tf32_lo(B)
    def timeit(fn, n=5):
        for _ in range(2): fn()
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
    ref = (A[:1, :256].double() @ B[:1].double())
    out = F.umma_gemm16(Ap, Bp, 0, a_tri=1)
    print("fp16x3 NN tri rel", rel(out[:1, :256], ref))
    out = F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo, a_tri=1)
    print("tf32x3 NN tri rel", rel(out[:1, :256], ref))
    for name, fn16, fn32, fl in (
        ("NN lower-tri", lambda: F.umma_gemm16(Ap, Bp, 0, a_tri=1), lambda: F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo, a_tri=1), L*M*M*N),
        ("NN full", lambda: F.umma_gemm16(Ap, Bp, 0), lambda: F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo), 2*L*M*M*N),
        ("NT full splitk4", lambda: F.umma_gemm16(Bp, Bp, 1, splitk=4), lambda: F.umma_gemm(B, B, 1, Alo=Blo, Blo=Blo, splitk=4), 2*L*M*M*N),
        ("NT tril splitk4", lambda: F.umma_gemm16(Bp, Bp, 1, d_tri=1, splitk=4), lambda: F.umma_gemm(B, B, 1, Alo=Blo, Blo=Blo, d_tri=1, splitk=4), L*M*M*N),
    ):
        t16, t32 = timeit(fn16), timeit(fn32)
        print(f"{name:18s} fp16x3 {t16:.3f} ms {fl/t16/1e9:7.1f} TF | tf32x3 {t32:.3f} ms {fl/t32/1e9:7.1f} TF")
    sd = torch.ones(L, device=dev)
    t = timeit(lambda: F.umma_gemm16(Ap, Bp, 0, a_tri=1, out_planes=True, out_scale=sd, want_amax=True))
    print(f"NN lower-tri -> planes+amax {t:.3f} ms")
