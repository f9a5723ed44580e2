"""Timing probe: formulations of the weight-gradient GEMM dW = dY' X of the PPO learner (K = 131072 rows), TF32.
usage: python tools/dw_gemm_probe.py   (needs a GPU)"""
import torch

torch.backends.cuda.matmul.allow_tf32 = True
dev = torch.device("cuda")


def bench(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


K = 131072
for (no, ni) in [(256, 640), (256, 256), (76, 256)]:
    gy, x = torch.randn(K, no, device=dev), torch.randn(K, ni, device=dev)
    ref = gy.t() @ x
    forms = {
        "gy.t() @ x": lambda: gy.t() @ x,
        "(x.t() @ gy).t()": lambda: (x.t() @ gy).t(),
        "copy-transposed gy, x": lambda: gy.t().contiguous() @ x.t().contiguous().t(),
    }
    for c in (4, 16, 64):
        forms[f"bmm {c} chunks + sum"] = (lambda c=c: torch.bmm(gy.view(c, K // c, no).transpose(1, 2), x.view(c, K // c, ni)).sum(0))
        forms[f"bmm {c} chunks (x' gy) + sum"] = (lambda c=c: torch.bmm(x.view(c, K // c, ni).transpose(1, 2), gy.view(c, K // c, no)).sum(0).t())
    for name, f in forms.items():
        err = float((f() - ref).abs().max() / ref.abs().max())
        print(f"[{no}x{ni}] {name:34s} {bench(f):8.1f} us  rel.err {err:.1e}", flush=True)
    for lib in ("cublaslt", "cublas"):
        torch.backends.cuda.preferred_blas_library(lib)
        print(f"[{no}x{ni}] gy.t() @ x with {lib:9s}          {bench(lambda: gy.t() @ x):8.1f} us", flush=True)
    torch.backends.cuda.preferred_blas_library("default")
