"""Role timeline (CTA 0) of the long-K project convs: plain and SE-gated."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
torch.manual_seed(0)


def run(name, fn, tiles=range(0, 5)):
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
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1", "bld0/epi2", "epi3")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))


for (c, co, hw) in ((960, 160, 14), (1536, 256, 7), (512, 128, 14)):
    x3 = torch.randn(512, hw, hw, c, device="cuda").bfloat16()
    w3 = (torch.randn(co, c, device="cuda") * 0.03).bfloat16()
    b3 = torch.zeros(co, device="cuda")
    r3 = torch.randn(512, hw, hw, co, device="cuda").bfloat16()
    g3 = torch.rand(512, c, device="cuda").bfloat16()
    run(f"conv1 {c}->{co} @{hw} project + residual", lambda: ops.conv_nhwc_bf16(x3, w3, 1, 1, bias=b3, act=None, residual=r3))
    run(f"conv1g {c}->{co} @{hw} gated project + residual", lambda: ops.conv1x1_gated(x3, g3, w3, bias=b3, act=None, residual=r3))
