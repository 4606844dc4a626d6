"""A/B a debug flag (ewvit_debug_set_flags) on the native backbone: python tools/bb_ab.py <flags> [frames]."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from torchvision.models import efficientnet_v2_s  # noqa: E402

from ewvit import engine  # noqa: E402
from ewvit._lib import load  # noqa: E402

lib = load()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
torch.manual_seed(0)
net = efficientnet_v2_s(weights=None).eval()
nb = engine.NativeEffNetV2(net.features, "cuda")
x = torch.randn(n, 3, 224, 224, device="cuda")
for fl in (0, flags, 0, flags):
    lib.ewvit_debug_set_flags(fl)
    for _ in range(3):
        nb.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        nb.forward(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"flags={fl}: {e0.elapsed_time(e1) / 10:.3f} ms per {n} frames")
lib.ewvit_debug_set_flags(0)
