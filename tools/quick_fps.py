"""Resident-input frames/s of the dynamic forward (512 frames), no stage timers (probe tool)."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from network.model import DeepfakeDetector
torch.manual_seed(42)
m = DeepfakeDetector(3, 128, batch_size=8).cuda().eval()
x = torch.randn(8, 64, 3, 224, 224, device="cuda")
with torch.no_grad():
    for _ in range(5):
        m(x, 8, "dynamic")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = m(x, 8, "dynamic")
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"{ms:.3f} ms/step, {512 / ms * 1e3:.0f} frames/s, logits finite: {bool(torch.isfinite(out['logits']).all())}")
