"""Calibration: what read-only and copy bandwidth do stock torch kernels reach on this GPU?"""
import torch, json
x = torch.empty(1 << 30, dtype=torch.int64, device="cuda").random_()   # 8 GiB
y = torch.empty(1 << 29, dtype=torch.int64, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: x.sum()); print(json.dumps({"op": "sum int64 8GiB", "ms": ms, "GBps": x.numel() * 8 / ms / 1e6}))
ms = t(lambda: (x == 12345).any()); print(json.dumps({"op": "eq.any 8GiB", "ms": ms, "GBps": x.numel() * 8 / ms / 1e6}))
xf = x.view(torch.float32)
ms = t(lambda: xf.max()); print(json.dumps({"op": "max f32 8GiB", "ms": ms, "GBps": x.numel() * 8 / ms / 1e6}))
ms = t(lambda: y.copy_(x[: 1 << 29])); print(json.dumps({"op": "copy 4GiB", "ms": ms, "GBps_rw": 2 * y.numel() * 8 / ms / 1e6}))
xb = x.view(torch.bfloat16); yb = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: yb.copy_(xb[: 1 << 30])); print(json.dumps({"op": "copy bf16 2GiB", "ms": ms, "GBps_rw": 2 * yb.numel() * 2 / ms / 1e6}))
