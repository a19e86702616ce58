"""Reference point for the pointwise GEMM: what does the library GEMM (torch.matmul -> cuBLASLt) reach on the
exact middle-flow shape (M = 256 tiles x 361 pixels, N = K = 728, bf16)?  Diagnostic only, never on the product path."""
import torch

def run(M, N, K, iters=30):
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    for _ in range(5):
        (a @ w.t())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        (a @ w.t())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")

for shape in [(92416, 728, 728), (92416, 768, 768), (92416, 1024, 728), (350464, 256, 256), (1401856, 128, 128), (25600, 1536, 1024), (25600, 2048, 1536)]:
    run(*shape)
