"""ncu target: one small-channel 3x3 conv (im2col builder path) + one 1x1 expand conv + one depthwise, 64 frames."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import engine, ops  # noqa: E402

n = 64
torch.manual_seed(0)
x = torch.randn(n, 112, 112, 24, device="cuda").bfloat16()
w = engine._w3x3_tapmajor_padded(torch.randn(24, 24, 3, 3) * 0.07).cuda()
b = torch.zeros(24, device="cuda")
x2 = torch.randn(n, 14, 14, 160, device="cuda").bfloat16()
w2 = (torch.randn(960, 160, device="cuda") * 0.08).bfloat16()
b2 = torch.zeros(960, device="cuda")
wd = torch.randn(9, 960, device="cuda") * 0.3
for _ in range(3):
    y = ops.conv_nhwc_bf16(x, w, 3, 1, bias=b, act="silu", residual=x)
    y2 = ops.conv_nhwc_bf16(x2, w2, 1, 1, bias=b2, act="silu")
    pooled = torch.empty(n, 960, device="cuda")
    y3 = ops.dwconv3x3(y2, wd, b2, 1, pooled=pooled)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(y3.float().abs().mean()))
