"""Correctness + timing of the flat 3x3 conv paths: per-tap tiles (debug flag 32) vs row-shared taps; flag 16 = base_offset variant."""
import os
import sys

import torch
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
torch.manual_seed(0)


def case(n, cin, cout):
    x = torch.zeros(n, 114, 114, cin, device="cuda", dtype=torch.bfloat16)
    x[:, 1:-1, 1:-1] = torch.randn(n, 112, 112, cin, device="cuda").bfloat16()
    w = (torch.randn(cout, 3, 3, cin, device="cuda") * (9 * cin) ** -0.5).bfloat16()
    sc, sh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.1
    y = torch.zeros(n, 114, 114, cout, device="cuda", dtype=torch.bfloat16)
    ref = None
    if n <= 4:
        xr = x[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
        ref = F.relu(F.conv2d(xr, w.permute(0, 3, 1, 2).float(), padding=1) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    for flags in (32, 64, 0):
        lib.ewvit_debug_set_flags(flags)
        y.zero_()
        ops.conv3x3_bf16(x, w, n, 112, 112, 1, True, sc, sh, True, y, 0, True)
        torch.cuda.synchronize()
        msg = ""
        if ref is not None:
            got = y[:, 1:-1, 1:-1].permute(0, 3, 1, 2).float()
            msg = f"max err {float((got - ref).abs().max()):.4f} (ref max {float(ref.abs().max()):.2f}) border {float(y[:, 0].abs().max())}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv3x3_bf16(x, w, n, 112, 112, 1, True, sc, sh, True, y, 0, True)
        e1.record()
        torch.cuda.synchronize()
        print(f"n={n} {cin}->{cout} flags={flags}: {e0.elapsed_time(e1) / 5:.3f} ms  {msg}")
    lib.ewvit_debug_set_flags(0)


case(2, 64, 128)
case(2, 384, 128)
case(256, 64, 128)
case(256, 384, 128)
