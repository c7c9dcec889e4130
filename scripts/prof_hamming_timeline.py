import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from ucfp_b200 import Context, Corpus, _ffi
n=int(float(sys.argv[1])); nq=int(sys.argv[2])
ctx = Context(0)
corpus = Corpus(ctx, _ffi.KIND_HAMMING64, n)
corpus.append_synthetic(0xC0DE, 0, n)
q = torch.from_numpy(oracle.fill_u64(nq, 0xBEEF).view(np.int64)).cuda()
ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda"); d = torch.empty((nq, 10), dtype=torch.int32, device="cuda")
for _ in range(3): corpus.scan_hamming(q, 10, ids, d)
torch.cuda.synchronize()
