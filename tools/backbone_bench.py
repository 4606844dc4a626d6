"""Per-layer CUDA-event timing of the native EfficientNetV2-S feature extractor (probe tool)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from torchvision.models import efficientnet_v2_s  # noqa: E402

from ewvit import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
net = efficientnet_v2_s(weights=None).eval()
nb = engine.NativeEffNetV2(net.features, "cuda")
x = torch.randn(n, 3, 224, 224, device="cuda")
for _ in range(3):
    nb.forward(x)
torch.cuda.synchronize()
engine.TIMER = engine.StageTimer()
engine.TIMER.per_layer = True
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    nb.forward(x)
e1.record()
torch.cuda.synchronize()
summ = engine.TIMER.summary_ms()
engine.TIMER = None
total = e0.elapsed_time(e1) / 5
kinds = {}
rows = []
for k, (ms, cnt) in sorted(summ.items()):
    kind = k.split(".")[-1]
    kinds[kind] = kinds.get(kind, 0.0) + ms
    op = nb.ops[int(k.split(".")[1])]
    shp = [tuple(t.shape) for t in op[1:] if hasattr(t, "shape")][:1]
    rows.append((k, round(ms, 4), shp))
print(json.dumps({"frames": n, "total_ms": total, "by_kind_ms": kinds}))
for r in rows:
    print(r)
