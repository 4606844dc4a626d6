"""How fast does the tcgen05 pipeline run with 128- vs 256-wide column tiles?  Long-K 1x1 convs (epilogue negligible)."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402

torch.manual_seed(0)
n, hw, k = 512, 14, 3456
x = torch.randn(n, hw, hw, k, device="cuda").bfloat16()
for cout in (128, 256, 512):
    w = (torch.randn(cout, k, device="cuda") * k ** -0.5).bfloat16()
    b = torch.zeros(cout, device="cuda")
    for _ in range(2):
        ops.conv_nhwc_bf16(x, w, 1, 1, bias=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_nhwc_bf16(x, w, 1, 1, bias=b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * n * hw * hw * k * cout
    print(f"1x1 conv K={k} N={cout}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
