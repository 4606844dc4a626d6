"""Role timeline (CTA 0) of the MWT hf-fusion conv 64 -> 128 @112x112 (row-shared taps, resident weights, CTA pairs)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
xh = torch.randn(n, 114, 114, cin, device="cuda").bfloat16()
wh = (torch.randn(128, 3, 3, cin, device="cuda") * 0.04).bfloat16()
yh = torch.empty(n, 114, 114, 128, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
fn = lambda: ops.conv3x3_bf16(xh, wh, n, 112, 112, 1, True, sc, sh, True, yh, 0, True)
for flags in (0, 128):
    lib.ewvit_debug_set_flags(flags)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * n * 112 * 112 * 128 * 9 * cin
    print(f"=== flags {flags} ({'1-CTA' if flags else 'CTA pairs'}): {ms * 1e3:.1f} us, {fl / ms / 1e9:.0f} TFLOP/s")
    buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
    lib.ewvit_debug_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.ewvit_debug_set_trace(None)
    t = buf.cpu().view(6, 64, 4)
    t0 = int(t[t > 0].min())
    for tile in range(8, 14):
        row = []
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))
lib.ewvit_debug_set_flags(0)
