"""ncu target: MWT head kernel on 64 frames (levels 1 and 3)."""
import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops
n = 64
torch.manual_seed(0)
w = torch.randn(3, 18, 3, 3, 3, device="cuda") * 0.2
sc, sh = torch.ones(54, device="cuda"), torch.zeros(54, device="cuda")
y = torch.zeros(n, 114, 114, 64, device="cuda", dtype=torch.bfloat16)
for hw in (112, 28):
    hf = torch.randn(n, 9, hw, hw, device="cuda")
    for _ in range(2):
        ops.mwt_head(hf, w, sc, sh, y, 112, 112)
torch.cuda.synchronize()
print("ok")
