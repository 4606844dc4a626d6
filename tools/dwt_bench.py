"""Time the fused 3-level Haar kernel on BASELINE.json configs[1] (256x3x224x224 fp32) with CUDA
events; print achieved algorithmic GB/s.  Used for quick GPU probes; bench.py reports the contract line."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.randn(n, 3, 224, 224, device="cuda")
out = ops.dwt3_haar(x)
alg_bytes = x.numel() * 4 + sum(v.numel() * 4 for v in out.values())
for _ in range(5):
    ops.dwt3_haar(x, out=out)
torch.cuda.synchronize()
iters = 50
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
ev[0].record()
for i in range(iters):
    ops.dwt3_haar(x, out=out)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
med = ts[len(ts) // 2]
# a plain device copy of the same byte volume, for reference on this very box
a = torch.empty(alg_bytes // 8, dtype=torch.float32, device="cuda")
b = torch.empty_like(a)
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
copy_ms = e0.elapsed_time(e1) / 20
print(json.dumps({"n": n, "alg_bytes": alg_bytes, "dwt3_ms_median": med, "dwt3_ms_min": ts[0],
                  "dwt3_GBps_median": alg_bytes / med / 1e6, "dwt3_GBps_best": alg_bytes / ts[0] / 1e6,
                  "copy_same_bytes_ms": copy_ms, "copy_GBps": alg_bytes / copy_ms / 1e6}))
