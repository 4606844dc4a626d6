"""Bring-up aid: per-role clock64 timeline of CTA 0 for a few representative tensor-core launches."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import engine, ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
n = 64
torch.manual_seed(0)


def run(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
    lib.ewvit_debug_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.ewvit_debug_set_trace(None)
    t = buf.cpu().view(6, 64, 4)
    t0 = int(t[t > 0].min())
    print(f"=== {name}")
    for tile in range(8, 13):
        row = []
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1", "bld0", "x")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))


x = torch.randn(256, 112, 112, 24, device="cuda").bfloat16()
w = engine._w3x3_tapmajor_padded(torch.randn(24, 24, 3, 3) * 0.07).cuda()
b = torch.zeros(24, device="cuda")
run("conv3 24->24 @112 (im2col, 4 kb/tile)", lambda: ops.conv_nhwc_bf16(x, w, 3, 1, bias=b, act="silu", residual=x))
x2 = torch.randn(512, 14, 14, 160, device="cuda").bfloat16()
w2 = (torch.randn(960, 160, device="cuda") * 0.08).bfloat16()
b2 = torch.zeros(960, device="cuda")
run("conv1 160->960 @14 (3 kb/tile)", lambda: ops.conv_nhwc_bf16(x2, w2, 1, 1, bias=b2, act="silu"))
xh = torch.randn(n, 114, 114, 64, device="cuda").bfloat16()
wh = (torch.randn(128, 3, 3, 64, device="cuda") * 0.04).bfloat16()
yh = torch.empty(n, 114, 114, 128, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
run("hf_fusion 64->128 @112 (9 kb/tile)", lambda: ops.conv3x3_bf16(xh, wh, n, 112, 112, 1, True, sc, sh, True, yh, 0, True))

x3 = torch.randn(512, 14, 14, 960, device="cuda").bfloat16()
w3 = (torch.randn(160, 960, device="cuda") * 0.03).bfloat16()
b3 = torch.zeros(160, device="cuda")
r3 = torch.randn(512, 14, 14, 160, device="cuda").bfloat16()
run("conv1 960->160 @14 project + residual (15 kb/tile)", lambda: ops.conv_nhwc_bf16(x3, w3, 1, 1, bias=b3, act=None, residual=r3))

x4 = torch.randn(256, 56, 56, 48, device="cuda").bfloat16()
w4 = engine._w3x3_tapmajor_padded(torch.randn(192, 48, 3, 3) * 0.05).cuda()
b4 = torch.zeros(192, device="cuda")
run("conv3 48->192 @56 (im2col, 7 kb/tile, bn 256)", lambda: ops.conv_nhwc_bf16(x4, w4, 3, 1, bias=b4, act="silu"))
x5 = torch.randn(256, 56, 56, 192, device="cuda").bfloat16()
w5 = (torch.randn(48, 192, device="cuda") * 0.07).bfloat16()
b5 = torch.zeros(48, device="cuda")
r5 = torch.randn(256, 56, 56, 48, device="cuda").bfloat16()
run("conv1 192->48 @56 project + residual (3 kb/tile)", lambda: ops.conv_nhwc_bf16(x5, w5, 1, 1, bias=b5, act=None, residual=r5))
