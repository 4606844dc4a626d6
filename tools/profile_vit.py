"""ncu target: the post-backbone SFE head on 512 frames, twice -- split-K patch_to_embedding (gemm_tc_kernel<EPI_PARTIAL> + reduce),
the ViT linears (gemm_tc_kernel<EPI_LINEAR>: qkv 1024x1536x512, out 1024x512x512, ff1 1024x2048x512, ff2 1024x512x2048, feat_map) and glue."""
import os
import sys

os.environ.setdefault("EWVIT_ALLOW_RANDOM_BACKBONE", "1")
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import engine  # noqa: E402
from network._native import load_architecture_config  # noqa: E402
from network.sfe import EfficientViT  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
cfg = load_architecture_config()
m = EfficientViT(cfg, channels=1280, selected_efficient_net=1, feat_dim=128, output_mode="feature_map")
sd = {k: v for k, v in m.state_dict().items() if not k.startswith("efficient_net.")}
run = engine.SfeRunner({k: v.cuda() for k, v in sd.items()}, cfg, backbone=None)
feat = (torch.randn(n, 62720, device="cuda") * 0.3).bfloat16()
pos = (torch.arange(n, device="cuda", dtype=torch.int32) % 64).contiguous()
for _ in range(2):
    y = run.head(feat, pos)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
