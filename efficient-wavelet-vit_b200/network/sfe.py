"""``network.sfe`` -- drop-in for the reference's spatial branch (network/sfe.py:12-173): EfficientNet
backbone -> one 62720-wide patch -> Linear -> [cls, patch] + pos -> 2 pre-norm Transformer blocks ->
``feat_map`` (feature_map mode) or ``mlp_head`` (cls mode).

Class names, constructor arguments, attribute names and state_dict keys follow the reference; eval-mode
CUDA calls run the native path (bf16 channels-last backbone + tcgen05 linears + fused token kernels).
"""
import os

import numpy as np
import torch
from torch import nn
from torchvision.models import efficientnet_v2_s

from ._native import NativeMixin

try:                                            # the reference's dependency, when the environment has it
    from efficientnet_pytorch import EfficientNet  # type: ignore
except ImportError:                             # same architecture / key names, random init (no network)
    from ._effnet_b0 import EfficientNet


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        return x + self.fn(x, **kwargs)


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_qkv = nn.Linear(dim, 3 * inner, bias=False)
        needs_projection = not (heads == 1 and dim_head == dim)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if needs_projection else nn.Identity()

    def forward(self, x):
        b, t, _ = x.shape
        q, k, v = (z.reshape(b, t, self.heads, -1).transpose(1, 2) for z in self.to_qkv(x).chunk(3, dim=-1))
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(b, t, -1)
        return self.to_out(out)


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                           PreNorm(dim, FeedForward(dim=dim, hidden_dim=mlp_dim, dropout=0))])
            for _ in range(depth)])

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return x


class EfficientViT(NativeMixin, nn.Module):
    def __init__(self, config, channels=512, selected_efficient_net=0, feat_dim=128, output_mode=None):
        super().__init__()
        self.output_mode = output_mode
        self.config = config
        m = config["model"]
        image_size, patch_size = m["image-size"], m["patch-size"]
        dim, depth, heads, dim_head = m["dim"], m["depth"], m["heads"], m["dim-head"]
        mlp_dim, emb_dim, num_classes = m["mlp-dim"], m["emb-dim"], m["num-classes"]
        assert image_size % patch_size == 0, "image dimensions must be divisible by the patch size"
        self.selected_efficient_net = selected_efficient_net

        if selected_efficient_net == 0:
            self.efficient_net = EfficientNet.from_pretrained("efficientnet-b0")
            if getattr(self.efficient_net, "_ewvit_random_init", False) and os.environ.get("EWVIT_ALLOW_RANDOM_BACKBONE", "0") != "1":
                _warn_random_backbone("EfficientNet-b0 (efficientnet_pytorch is not installed; built-in architecture copy)",
                                      ImportError("no pretrained source"))
        else:
            self.efficient_net = efficientnet_v2_s(weights=_v2s_weights())
            self.efficient_net.classifier = nn.Identity()
        # first six backbone tensors stay frozen, the rest train (sfe.py:115-119)
        for index, (_, param) in enumerate(self.efficient_net.named_parameters()):
            param.requires_grad = index > 5

        self.patch_size = patch_size
        patch_dim = channels * patch_size ** 2
        self.pos_embedding = nn.Parameter(torch.randn(emb_dim, 1, dim))
        self.patch_to_embedding = nn.Linear(patch_dim, dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(m["emb-dropout"])
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, m["dropout"])
        self.to_cls_token = nn.Identity()
        self.mlp_head = nn.Sequential(nn.Linear(dim, mlp_dim), nn.ReLU(), nn.Linear(mlp_dim, num_classes))
        self.feat_map = nn.Sequential(nn.Linear(dim, feat_dim), nn.ReLU())

    # ---- PyTorch composition (training / hooks); mirrors sfe.py:145-173
    def _backbone(self, img):
        if self.selected_efficient_net == 0:
            return self.efficient_net.extract_features(img)
        return self.efficient_net.features(img)

    def _forward_torch(self, img):
        p = self.patch_size
        x = self._backbone(img)
        b, c, hh, ww = x.shape
        y = x.reshape(b, c, hh // p, p, ww // p, p).permute(0, 2, 4, 3, 5, 1).reshape(b, (hh // p) * (ww // p), p * p * c)
        y = self.patch_to_embedding(y)
        x = torch.cat((self.cls_token.expand(b, -1, -1), y), dim=1)
        x = x + self.pos_embedding[0:b]                  # raises for b > emb-dim, like the reference
        x = self.transformer(self.dropout(x))
        if self.output_mode == "cls":
            return self.mlp_head(self.to_cls_token(x[:, 0]))
        side = int(np.sqrt(x.shape[1] - 1))
        t = self.feat_map(x[:, 1:])
        return t.reshape(b, side, side, -1).permute(0, 3, 1, 2)

    # ---- native path
    def _build_runner(self):
        from ewvit.engine import SfeRunner, make_backbone
        feats = self.efficient_net.features if self.selected_efficient_net != 0 else self.efficient_net
        dev = self.pos_embedding.device
        backbone = make_backbone(feats, dev, v2s=self.selected_efficient_net != 0)
        sd = {k: v for k, v in self.state_dict().items() if not k.startswith("efficient_net.")}
        return SfeRunner(sd, self.config, backbone, "cls" if self.output_mode == "cls" else "feature_map")

    def forward(self, img, mask=None):
        if self._use_native(img):
            from ewvit.engine import check_chunk_limit
            n = img.shape[0]
            check_chunk_limit(n, 1, 1, self.pos_embedding.shape[0])
            runner = self._native_runner(self._build_runner)
            pos = torch.arange(n, dtype=torch.int32, device=img.device)
            out = runner.forward(img.float().contiguous(), pos)
            if self.output_mode == "cls":
                return out
            return out.clone().view(n, -1, 1, 1)
        return self._forward_torch(img)


def _v2s_weights():
    """The reference builds ``efficientnet_v2_s(weights=EfficientNet_V2_S_Weights.IMAGENET1K_V1)`` (sfe.py:111-112) and
    freezes its first six tensors.  Same here whenever the checkpoint is in the torch hub cache or can be downloaded;
    when it cannot (no network), warn loudly -- the backbone is then RANDOMLY initialised with a frozen random stem,
    which is fine for loading a trained checkpoint or for benchmarks but not for training from scratch.
    ``EWVIT_ALLOW_RANDOM_BACKBONE=1`` (tests, benchmarks) skips the attempt and the warning."""
    if os.environ.get("EWVIT_ALLOW_RANDOM_BACKBONE", "0") == "1":
        return None
    from torchvision.models import EfficientNet_V2_S_Weights
    weights = EfficientNet_V2_S_Weights.IMAGENET1K_V1
    import socket
    old = socket.getdefaulttimeout()
    try:
        socket.setdefaulttimeout(10)
        weights.get_state_dict(progress=False, check_hash=False)       # hub cache hit, or download
        return weights
    except Exception as e:                                             # noqa: BLE001 (URLError, OSError, hash errors ...)
        _warn_random_backbone("EfficientNetV2-S (torchvision IMAGENET1K_V1)", e)
        return None
    finally:
        socket.setdefaulttimeout(old)


def _warn_random_backbone(what, err):
    import warnings
    warnings.warn(f"{what} ImageNet weights could not be loaded ({type(err).__name__}: {err}); the reference starts from "
                  "them (network/sfe.py:109-112).  This backbone is RANDOMLY initialised and its first six tensors are "
                  "frozen as in the reference: load a trained checkpoint before use, or put the weights in the torch hub "
                  "cache.  Set EWVIT_ALLOW_RANDOM_BACKBONE=1 to silence this.", stacklevel=3)
