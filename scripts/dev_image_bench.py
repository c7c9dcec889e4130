"""Developer timing for the image multi-hash kernels (not the contract bench)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from ucfp_b200 import Context, _ffi
ctx = Context(0)
for (w, h, n) in [(256, 256, 16384), (1024, 1024, 2048), (640, 480, 4096), (1920, 1080, 1024)]:
    px = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
    out = torch.zeros((n, 51), dtype=torch.int64, device="cuda")
    for _ in range(2): ctx.image_hash_uniform(px, n, w, h, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps): ctx.image_hash_uniform(px, n, w, h, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"w": w, "h": h, "n": n, "ms": round(ms, 3), "img_per_s": round(n / ms * 1e3), "GBps": round(n * w * h * 3 / ms / 1e6, 1)}), flush=True)
