"""One fixed image-hash batch for ncu (developer tool)."""
import sys
import torch
sys.path.insert(0, ".")
from ucfp_b200 import Context
w, h, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = Context(0)
px = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
out = torch.zeros((n, 51), dtype=torch.int64, device="cuda")
for _ in range(3):
    ctx.image_hash_uniform(px, n, w, h, out=out)
torch.cuda.synchronize()
print("ok")
