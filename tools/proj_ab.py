"""Project-conv A/B (512 frames): residual on/off, CTA pairs on/off (debug flag 128) -- attributes the epilogue's share."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402
lib = load()


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for (c, co, hw) in ((960, 160, 14), (1536, 256, 7), (512, 128, 14), (160, 960, 14), (256, 1536, 7)):
    xs = [torch.randn(512, hw, hw, c, device="cuda").bfloat16() for _ in range(3)]
    w = (torch.randn(co, c, device="cuda") * 0.03).bfloat16()
    b = torch.zeros(co, device="cuda")
    r = torch.randn(512, hw, hw, co, device="cuda").bfloat16()
    g = torch.rand(512, c, device="cuda").bfloat16()
    y = torch.empty(512, hw, hw, co, device="cuda", dtype=torch.bfloat16)
    i = [0]
    def nxt():
        i[0] += 1
        return xs[i[0] % 3]
    row = []
    for flags in (0, 128):
        lib.ewvit_debug_set_flags(flags)
        if co < c:
            row.append(f"[{'pair' if flags == 0 else '1cta'}] plain+res {t(lambda: ops.conv_nhwc_bf16(nxt(), w, 1, 1, bias=b, residual=r, out=y)):6.1f}  "
                       f"plain {t(lambda: ops.conv_nhwc_bf16(nxt(), w, 1, 1, bias=b, out=y)):6.1f}  "
                       f"gated+res {t(lambda: ops.conv1x1_gated(nxt(), g, w, bias=b, residual=r, out=y)):6.1f}  "
                       f"gated {t(lambda: ops.conv1x1_gated(nxt(), g, w, bias=b, out=y)):6.1f}")
        else:
            row.append(f"[{'pair' if flags == 0 else '1cta'}] expand+silu {t(lambda: ops.conv_nhwc_bf16(nxt(), w * 0.5, 1, 1, bias=b, act='silu_h', out=y)):6.1f}  "
                       f"expand linear {t(lambda: ops.conv_nhwc_bf16(nxt(), w, 1, 1, bias=b, out=y)):6.1f}")
    lib.ewvit_debug_set_flags(0)
    byt = 512 * hw * hw * (c + co) * 2
    print(f"{c:5d}->{co:5d} @{hw:2d} (HBM floor {byt / 6.5456e6:5.1f} us)  " + "   ".join(row))
