"""Timing of the tensor-core MWT head pieces (upsample per level, block-diagonal conv) + role trace of the conv."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
n = 512
torch.manual_seed(0)


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us")


up = torch.zeros(n, 114, 114, 16, device="cuda", dtype=torch.bfloat16)
y = torch.zeros(n, 114, 114, 64, device="cuda", dtype=torch.bfloat16)
for hin in (112, 56, 28):
    hf = torch.randn(n, 9, hin, hin, device="cuda")
    timeit(f"upsample {hin}->112", lambda: ops.mwt_upsample(hf, up, 112, 112))
w = (torch.randn(64, 192, device="cuda") * 0.1).bfloat16()
sc, sh = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
timeit("head conv", lambda: ops.mwt_head_conv(up, w, sc, sh, y, 112, 112))
buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
lib.ewvit_debug_set_trace(buf.data_ptr())
ops.mwt_head_conv(up, w, sc, sh, y, 112, 112)
torch.cuda.synchronize()
lib.ewvit_debug_set_trace(None)
t = buf.cpu().view(6, 64, 4)
t0 = int(t[t > 0].min())
for tile in range(8, 14):
    row = []
    for role, nm in enumerate(("tma", "mma", "epi0", "epi1", "r4", "r5")):
        v = t[role, tile]
        if int(v.max()) == 0:
            continue
        row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
    print(f"tile {tile:2d}  " + "  ".join(row))
