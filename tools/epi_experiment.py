"""Bring-up experiment: time one wide 1x1 expand conv with parts of the epilogue switched off."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops
from ewvit._lib import load
lib = load()
x = torch.randn(512, 14, 14, 160, device="cuda").bfloat16()
w = (torch.randn(960, 160, device="cuda") * 0.08).bfloat16()
b = torch.zeros(960, device="cuda")
out = torch.empty(512, 14, 14, 960, device="cuda", dtype=torch.bfloat16)
def t(flags, act):
    lib.ewvit_debug_set_flags(flags)
    for _ in range(3): ops.conv_nhwc_bf16(x, w, 1, 1, bias=b, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.conv_nhwc_bf16(x, w, 1, 1, bias=b, act=act, out=out)
    e1.record(); torch.cuda.synchronize()
    lib.ewvit_debug_set_flags(0)
    return e0.elapsed_time(e1) / 20
print("full silu        :", t(0, "silu"))
print("no act           :", t(0, None))
print("silu, no stores  :", t(1, "silu"))
print("no act, no stores:", t(1, None))
