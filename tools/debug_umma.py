import sys, torch
sys.path.insert(0, '.')
from gpzoo_b200 import functional as F
torch.manual_seed(0)
dev = 'cuda'
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
for bk in (1, 0):
    for (m, n, k) in ((128, 256, 16), (128, 256, 64), (256, 512, 128)):
        A = torch.randn(1, m, k, device=dev)
        B = torch.randn((1, n, k) if bk else (1, k, n), device=dev)
        ref = A.double() @ (B.double().transpose(1, 2) if bk else B.double())
        for nt in (1, 3):
            out = F.umma_gemm(A, B, bk, n_terms=nt)
            torch.cuda.synchronize()
            print(f"bk={bk} m{m} n{n} k{k} terms={nt}: rel={rel(out, ref):.3e} nz={int((out != 0).sum())}/{out.numel()} "
                  f"out[0,0,:4]={out[0,0,:4].tolist()} ref={ref[0,0,:4].tolist()}")
# structured probe: A = identity-like, to read back B layout
m, n, k = 128, 256, 16
A = torch.zeros(1, m, k, device=dev); A[0, torch.arange(16), torch.arange(16)] = 1.0
for bk in (1, 0):
    B = (torch.arange(n * k, device=dev, dtype=torch.float32).reshape(1, n, k) if bk else
         torch.arange(n * k, device=dev, dtype=torch.float32).reshape(1, k, n))
    out = F.umma_gemm(A, B, bk, n_terms=1)
    ref = A @ (B.transpose(1, 2) if bk else B)
    print("probe bk", bk, "match", bool(torch.equal(out, ref)), out[0, :3, :6].tolist(), ref[0, :3, :6].tolist())
