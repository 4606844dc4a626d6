"""Short, deterministic target for ncu: the native MWT branch (DWT -> head -> tcgen05 convs) on 64 frames, twice."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from network.mwt import MWT  # noqa: E402

torch.manual_seed(0)
m = MWT().cuda().eval()
x = torch.randn(int(sys.argv[1]) if len(sys.argv) > 1 else 64, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m(x)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
