"""Role timeline (CTA 0) of the MWT head conv 9(16) -> 54(64) @112x112 (K = 16 taps on 32B-swizzled windows)."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402
lib = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
up = torch.randn(n, 114, 114, 16, device="cuda").bfloat16()
w = (torch.randn(64, 144, device="cuda") * 0.1).bfloat16()
y = torch.empty(n, 114, 114, 64, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
fn = lambda: ops.mwt_head_conv(up, w, sc, sh, y, 112, 112)
for flags in (0, 512):
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
    print(f"=== flags {flags}: {ms * 1e3:.1f} us ({(up.numel() + y.numel()) * 2 / ms / 1e6:.0f} GB/s)")
    buf = torch.zeros(6 * 64 * 4, dtype=torch.int64, device="cuda")
    lib.ewvit_debug_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.ewvit_debug_set_trace(None)
    t = buf.cpu().view(6, 64, 4)
    t0 = int(t[t > 0].min())
    for tile in range(20, 26):
        row = []
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))
lib.ewvit_debug_set_flags(0)
