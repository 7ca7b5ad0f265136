"""cuBLAS 8192^3 matmul throughput per dtype on this GPU (the denominators next to MEASURED_PEAKS.json's bf16 figure)."""
import torch
dev = 'cuda'
def tf(dtype, allow_tf32=False, n=8192, reps=20):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3): a @ b
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): a @ b
    e1.record(); torch.cuda.synchronize()
    return 2 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
print("bf16  %.0f TFLOP/s" % tf(torch.bfloat16))
print("fp16  %.0f TFLOP/s" % tf(torch.float16))
print("tf32  %.0f TFLOP/s" % tf(torch.float32, True))
print("fp32  %.0f TFLOP/s" % tf(torch.float32, False, reps=5))
print("fp64  %.0f TFLOP/s" % tf(torch.float64, False, n=4096, reps=5))
