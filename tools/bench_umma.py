import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
torch.manual_seed(0)
dev = 'cuda'
L, M, N = 10, 1024, 32768
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
A = torch.tril(torch.randn(L, M, M, device=dev)); B = torch.randn(L, M, N, device=dev)
Alo, Blo = F.tf32_lo(A), F.tf32_lo(B)
ref = (A[:1, :256].double() @ B[:1].double())
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
out = F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo, a_tri=1)
print("NN tri rel", rel(out[:1, :256], ref))
t = timeit(lambda: F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo, a_tri=1))
print(f"NN lower-tri  {t:.3f} ms  {L*M*M*N/t/1e9:.1f} TFLOP/s (tri flops)")
t = timeit(lambda: F.umma_gemm(A, B, 0, Alo=Alo, Blo=Blo))
print(f"NN full       {t:.3f} ms  {2*L*M*M*N/t/1e9:.1f} TFLOP/s")
Bt = torch.randn(L, M, N, device=dev); Btlo = F.tf32_lo(Bt)
t = timeit(lambda: F.umma_gemm(B, Bt, 1, Alo=Blo, Blo=Btlo, d_tri=1, splitk=4))
print(f"NT tril splitk4 {t:.3f} ms")
t = timeit(lambda: F.umma_gemm(B, Bt, 1, Alo=Blo, Blo=Btlo, splitk=4))
print(f"NT full splitk4 {t:.3f} ms  {2*L*M*M*N/t/1e9:.1f} TFLOP/s")
