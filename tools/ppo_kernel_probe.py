"""Timing of the fused tanh-Normal kernels (csrc/bt_ppo.cu) at the learner's minibatch shape against their algorithmic bytes.
usage: python tools/ppo_kernel_probe.py   (needs a GPU)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from brax_tracking_b200 import native

B, T, A = 8192, 16, 38
dev = torch.device("cuda")
logits = torch.randn(B, T, 2 * A, device=dev)
raw = torch.randn(B, T, A, device=dev)
noise = torch.randn(T, B, A, device=dev).transpose(0, 1)
glp, gent = torch.randn(T, B, device=dev), torch.randn(T, B, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(f, n=20):
    f(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()                                                        # L2 flush between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


rows = B * T
fwd_ms = timed(lambda: native.ppo_tanh_normal(logits, raw, noise))
bwd_ms = timed(lambda: native.ppo_tanh_normal(logits, raw, noise, grads=(glp, gent)))
fwd_b, bwd_b = rows * (4 * A * 4 + 8), rows * (4 * A * 4 + 8 + 2 * A * 4)
peak = 6545.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"rows": rows, "A": A, "fwd_us": round(fwd_ms * 1e3, 1), "fwd_GBs": round(fwd_b / fwd_ms / 1e6, 1), "fwd_frac": round(fwd_b / fwd_ms / 1e6 / peak, 3),
                  "bwd_us": round(bwd_ms * 1e3, 1), "bwd_GBs": round(bwd_b / bwd_ms / 1e6, 1), "bwd_frac": round(bwd_b / bwd_ms / 1e6 / peak, 3),
                  "algorithmic_bytes": {"fwd": fwd_b, "bwd": bwd_b}, "peak_GBs": peak, "l2": "flushed between timed launches"}))
