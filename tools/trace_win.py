"""Role timeline of the overlapping-window 3x3 convs (24->24 @112, 48->192 @56)."""
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import engine, ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
torch.manual_seed(0)


def run(name, fn, tiles=range(8, 13)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
    lib.ewvit_debug_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.ewvit_debug_set_trace(None)
    t = buf.cpu().view(6, 64, 4)
    t0 = int(t[t > 0].min())
    print(f"=== {name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
    for tile in tiles:
        row = []
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1", "epi2", "epi3")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))


n = 512
for cin, cout, hw, res in ((24, 24, 112, True), (48, 192, 56, False)):
    x = torch.zeros(n, hw + 2, hw + 2, cin, device="cuda", dtype=torch.bfloat16)
    x[:, 1:-1, 1:-1] = torch.randn(n, hw, hw, cin, device="cuda").bfloat16()
    w = engine._w3x3_window_packed(torch.randn(cout, cin, 3, 3) * 0.05).cuda()
    b = torch.zeros(cout, device="cuda")
    run(f"window conv3 {cin}->{cout} @{hw}", lambda: ops.conv_nhwc_bf16_ex(x, w, 3, 1, cin, bias=b, act="silu", residual=x if res else None,
                                                                           in_padded=True, out_padded=True))
