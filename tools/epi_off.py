"""Diagnosis: MWT 3x3 convs with the epilogue switched off (debug flag 512) -- how fast is the operand feed + MMA alone?"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
torch.manual_seed(0)
n = 256
for cin in (64, 384):
    x = torch.zeros(n, 114, 114, cin, device="cuda", dtype=torch.bfloat16)
    x[:, 1:-1, 1:-1] = torch.randn(n, 112, 112, cin, device="cuda").bfloat16()
    w = (torch.randn(128, 3, 3, cin, device="cuda") * (9 * cin) ** -0.5).bfloat16()
    sc, sh = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    y = torch.zeros(n, 114, 114, 128, device="cuda", dtype=torch.bfloat16)
    for flags in (0, 512, 0, 512):
        lib.ewvit_debug_set_flags(flags)
        for _ in range(2):
            ops.conv3x3_bf16(x, w, n, 112, 112, 1, True, sc, sh, True, y, 0, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv3x3_bf16(x, w, n, 112, 112, 1, True, sc, sh, True, y, 0, True)
        e1.record()
        torch.cuda.synchronize()
        print(f"{cin}->128 n={n} flags={flags}: {e0.elapsed_time(e1) / 5:.3f} ms")
lib.ewvit_debug_set_flags(0)
