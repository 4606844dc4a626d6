"""CPU fp32 restatement of the EWViT per-frame forward path.  TEST INFRASTRUCTURE ONLY.

Functional (state_dict in, tensors out) restatement of the reference's eval-mode
forward, used as the parity checker for the CUDA path and as the timed CPU baseline.
Each function cites the reference lines it follows.  It is validated against the
UNMODIFIED reference modules by ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` exists); the resulting fixtures are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py`` on any box.

The Haar DWT is third-party arithmetic (see ``oracle/haar.py``: PARITY UNPINNED).
The EfficientNetV2-S backbone is third-party arithmetic too (torchvision, present on
both boxes); the oracle calls torchvision's own module for it (``network/sfe.py:111-113``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .haar import haar_dwt2

SD = Dict[str, torch.Tensor]

# config/architecture.yaml:1-11 (field names kept)
DEFAULT_CONFIG = {
    "model": {
        "image-size": 224, "patch-size": 7, "num-classes": 1, "dim": 512, "depth": 2,
        "dim-head": 64, "heads": 8, "mlp-dim": 2048, "emb-dim": 64,
        "dropout": 0.15, "emb-dropout": 0.15,
    }
}

BN_EPS = 1e-5   # nn.BatchNorm2d default, network/mwt.py:34
LN_EPS = 1e-5   # nn.LayerNorm default, network/sfe.py:23


def _conv_bn_relu(x, sd: SD, conv: str, bn: str, stride=1):
    """Conv3x3(pad 1) -> BatchNorm(eval) -> ReLU, e.g. network/mwt.py:33-36."""
    y = F.conv2d(x, sd[conv + "weight"], sd[conv + "bias"], stride=stride, padding=1)
    y = F.batch_norm(y, sd[bn + "running_mean"], sd[bn + "running_var"],
                     sd[bn + "weight"], sd[bn + "bias"], False, 0.0, BN_EPS)
    return F.relu(y)


# --------------------------------------------------------------------------- MWT
def mwt_wavelet_transform(sd: SD, p: str, x, target_size, levels=3):
    """network/mwt.py:74-90.  Returns (ll, hf_compressed, hf9) -- hf9 is the upsampled
    9-channel high-frequency map fed to the separate convs (kept for kernel tests)."""
    n, c, h, w = x.shape
    ll, yh = haar_dwt2(x)                      # mwt.py:76
    hf = yh.reshape(n, 3 * c, yh.shape[-2], yh.shape[-1])   # mwt.py:77: colour-major, subband-minor
    if levels > 1:
        hf = F.interpolate(hf, size=target_size, mode="bilinear")   # mwt.py:79-81
    parts = []
    for i in range(3):                         # mwt.py:84-86: group i = channels [i*C, (i+1)*C)
        q = f"{p}hf_conv.seperate.{i}."
        parts.append(_conv_bn_relu(hf[:, i * c:(i + 1) * c], sd, q + "0.", q + "1."))
    cat = torch.cat(parts, dim=1)
    q = f"{p}hf_conv.fusion."
    return ll, _conv_bn_relu(cat, sd, q + "0.", q + "1."), hf


def mwt_forward(sd: SD, p: str, x, levels=3, return_intermediates=False):
    """network/mwt.py:92-119.  x [N,3,H,W] -> [N,dim,1,1]."""
    n, c, h, w = x.shape
    target = (h // 2, w // 2)
    cur = x
    feats = []
    inter = {}
    for lvl in range(levels):                  # mwt.py:107-111
        ll, hf, hf9 = mwt_wavelet_transform(sd, p, cur, target, levels)
        feats.append(hf)
        if return_intermediates:
            inter[f"hf9_{lvl}"] = hf9
            inter[f"hfc_{lvl}"] = hf
        cur = ll
    ms = torch.cat(feats, dim=1)               # mwt.py:113
    y = _conv_bn_relu(ms, sd, p + "multiscale_fusion.0.", p + "multiscale_fusion.1.")   # :114
    if return_intermediates:
        inter["multiscale"] = y
    y = _conv_bn_relu(y, sd, p + "freq_conv.0.", p + "freq_conv.1.", stride=2)          # :116
    if return_intermediates:
        inter["freq_conv"] = y
    y = F.max_pool2d(y, 2, 2)                  # freq_pool.0, mwt.py:39
    y = _conv_bn_relu(y, sd, p + "freq_pool.1.", p + "freq_pool.2.", stride=2)          # :40-42
    y = F.adaptive_avg_pool2d(y, 1)            # :43
    if return_intermediates:
        return y, inter
    return y


# --------------------------------------------------------------------------- SFE
_BACKBONE_CACHE: dict = {}


def backbone_v2s_features(sd: SD, p: str, x):
    """torchvision efficientnet_v2_s(...).features, network/sfe.py:111-113,150. p = "<...>.efficient_net."."""
    from torchvision.models import efficientnet_v2_s
    key = id(sd), p, str(x.device)
    net = _BACKBONE_CACHE.get(key)
    if net is None:
        net = efficientnet_v2_s(weights=None)
        net.classifier = torch.nn.Identity()
        sub = {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
        net.load_state_dict(sub, strict=True)
        net.eval().to(x.device)
        _BACKBONE_CACHE.clear()
        _BACKBONE_CACHE[key] = net
    with torch.no_grad():
        return net.features(x)


def _layer_norm(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], LN_EPS)


def vit_attention(sd: SD, p: str, x, heads: int, dim_head: int):
    """network/sfe.py:42-70 (Attention) on already-normalised tokens x [N,T,D]."""
    n, t, _ = x.shape
    qkv = x @ sd[p + "to_qkv.weight"].t()
    q, k, v = qkv.chunk(3, dim=-1)
    sh = lambda z: z.reshape(n, t, heads, dim_head).permute(0, 2, 1, 3)
    q, k, v = sh(q), sh(k), sh(v)
    dots = (q @ k.transpose(-1, -2)) * (dim_head ** -0.5)
    attn = dots.softmax(dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(n, t, heads * dim_head)
    return out @ sd[p + "to_out.0.weight"].t() + sd[p + "to_out.0.bias"]


def vit_tokens_from_features(sd: SD, p: str, feat, cfg=DEFAULT_CONFIG):
    """network/sfe.py:153-159: patch flatten (NHWC order), patch_to_embedding, cls, pos."""
    n = feat.shape[0]
    ps = cfg["model"]["patch-size"]
    assert feat.shape[-1] == ps and feat.shape[-2] == ps, "one patch per frame at the shipped config"
    y = feat.permute(0, 2, 3, 1).reshape(n, 1, -1)           # (p1 p2 c)
    y = y @ sd[p + "patch_to_embedding.weight"].t() + sd[p + "patch_to_embedding.bias"]
    x = torch.cat([sd[p + "cls_token"].expand(n, -1, -1), y], dim=1)
    pos = sd[p + "pos_embedding"][0:n]                        # sfe.py:158-159
    if pos.shape[0] != n:
        raise RuntimeError(
            f"The size of tensor a ({n}) must match the size of tensor b ({pos.shape[0]}) "
            "at non-singleton dimension 0")
    return x + pos


def vit_transformer(sd: SD, p: str, x, cfg=DEFAULT_CONFIG):
    """network/sfe.py:72-85."""
    m = cfg["model"]
    for l in range(m["depth"]):
        q = f"{p}transformer.layers.{l}."
        x = vit_attention(sd, q + "0.fn.", _layer_norm(x, sd, q + "0.norm."), m["heads"], m["dim-head"]) + x
        hdn = F.gelu(_layer_norm(x, sd, q + "1.norm.") @ sd[q + "1.fn.net.0.weight"].t() + sd[q + "1.fn.net.0.bias"])
        x = hdn @ sd[q + "1.fn.net.3.weight"].t() + sd[q + "1.fn.net.3.bias"] + x
    return x


def sfe_head(sd: SD, p: str, feat, cfg=DEFAULT_CONFIG, output_mode="feature_map"):
    """Everything after the backbone, network/sfe.py:153-173."""
    x = vit_transformer(sd, p, vit_tokens_from_features(sd, p, feat, cfg), cfg)
    if output_mode == "cls":                                  # sfe.py:163-166
        h = F.relu(x[:, 0] @ sd[p + "mlp_head.0.weight"].t() + sd[p + "mlp_head.0.bias"])
        return h @ sd[p + "mlp_head.2.weight"].t() + sd[p + "mlp_head.2.bias"]
    t = F.relu(x[:, 1:] @ sd[p + "feat_map.0.weight"].t() + sd[p + "feat_map.0.bias"])  # sfe.py:171
    n, tk, d = t.shape
    s = int(math.isqrt(tk))
    return t.reshape(n, s, s, d).permute(0, 3, 1, 2)


def sfe_forward(sd: SD, p: str, x, cfg=DEFAULT_CONFIG, output_mode="feature_map"):
    """network/sfe.py:145-173 with the V2-S backbone (selected_efficient_net=1)."""
    return sfe_head(sd, p, backbone_v2s_features(sd, p + "efficient_net.", x), cfg, output_mode)


def sfe_b0_forward(sd: SD, p: str, x, cfg=DEFAULT_CONFIG, output_mode="feature_map"):
    """network/sfe.py:145-173 with the EfficientNet-b0 backbone (selected_efficient_net=0, sfe.py:109,148):
    the standalone ``sfe`` (feature map) and ``sfe_cls`` (``output_mode='cls'``) branches of model.py:38-51."""
    from .effnet_b0 import extract_features
    return sfe_head(sd, p, extract_features(sd, p + "efficient_net.", x), cfg, output_mode)


# --------------------------------------------------------------------------- DAMA
def cross_attention(sd: SD, p: str, xn, ctx, heads: int):
    """network/dama.py:33-53 with kv_include_self=True: keys/values = cat(xn, ctx)."""
    n, t, d = xn.shape
    dh = sd[p + "to_q.weight"].shape[0] // heads
    kv_in = torch.cat([xn, ctx], dim=1)
    q = xn @ sd[p + "to_q.weight"].t()
    k, v = (kv_in @ sd[p + "to_kv.weight"].t()).chunk(2, dim=-1)
    sh = lambda z: z.reshape(n, z.shape[1], heads, dh).permute(0, 2, 1, 3)
    q, k, v = sh(q), sh(k), sh(v)
    attn = ((q @ k.transpose(-1, -2)) * (dh ** -0.5)).softmax(dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(n, t, heads * dh)
    return out @ sd[p + "to_out.0.weight"].t() + sd[p + "to_out.0.bias"]


def bidirectional_cross(sd: SD, p: str, s, f, heads=4, depth=2):
    """network/dama.py:68-78."""
    for l in range(depth):
        q = f"{p}layers.{l}."
        s = s + cross_attention(sd, q + "1.", _layer_norm(s, sd, q + "0."), f, heads)
        f = f + cross_attention(sd, q + "3.", _layer_norm(f, sd, q + "2."), s, heads)
    return s, f


def dama_fuse(sd: SD, p: str, space, freq, heads=4):
    """network/dama.py:143-169 after the two branches. space/freq [N,D,1,1]."""
    n, d, hh, ww = space.shape
    s = space.flatten(2).transpose(1, 2)
    f = freq.flatten(2).transpose(1, 2)
    s, f = bidirectional_cross(sd, p + "cross_att.", s, f, heads)
    space = s.transpose(1, 2).reshape(n, d, hh, ww)
    freq = f.transpose(1, 2).reshape(n, d, hh, ww)
    cat = torch.cat([space, freq], dim=1)
    fused = _conv_bn_relu(cat, sd, p + "fusion_gate.0.", p + "fusion_gate.1.")   # dama.py:124-128,153
    g = F.adaptive_avg_pool2d(cat, 1).flatten(1)                                 # gate_net.0/.1
    g = F.relu(g @ sd[p + "gate_net.2.weight"].t() + sd[p + "gate_net.2.bias"])
    g = (g @ sd[p + "gate_net.5.weight"].t() + sd[p + "gate_net.5.bias"]).softmax(dim=1)
    w = (g[:, 0].view(n, 1, 1, 1) * space + g[:, 1].view(n, 1, 1, 1) * freq
         + g[:, 2].view(n, 1, 1, 1) * fused)                                     # dama.py:159-163
    return {"fused": w.mean(dim=[2, 3]), "space": space.mean(dim=[2, 3]), "freq": freq.mean(dim=[2, 3])}


def dama_process_frame(sd: SD, p: str, frames, cfg=DEFAULT_CONFIG, levels=3, heads=4):
    """network/dama.py:130-169. frames [N,3,H,W]."""
    space = sfe_forward(sd, p + "sfe.", frames, cfg)
    freq = mwt_forward(sd, p + "mwt.", frames, levels)
    return dama_fuse(sd, p, space, freq, heads)


def dama_forward(sd: SD, p: str, x, batch_size=16, cfg=DEFAULT_CONFIG):
    """network/dama.py:171-206: serial chunk loop over K, per-video mean."""
    b, k = x.shape[:2]
    dim = sd[p + "gate_net.2.weight"].shape[1] // 2
    acc = {n: torch.zeros(b, dim, device=x.device) for n in ("fused", "space", "freq")}
    for s in range(0, k, batch_size):
        e = min(s + batch_size, k)
        out = dama_process_frame(sd, p, x[:, s:e].flatten(0, 1), cfg)
        for n in acc:
            acc[n] += out[n].view(b, -1, dim).sum(dim=1)
    return {n: v / k for n, v in acc.items()}


# --------------------------------------------------------------------------- detector
def classifier(sd: SD, feats):
    """network/model.py:62-68 (eval: dropout off)."""
    h = F.relu(feats @ sd["classifier.0.weight"].t() + sd["classifier.0.bias"])
    return h @ sd["classifier.3.weight"].t() + sd["classifier.3.bias"]


def detector_forward(sd: SD, x, batch_size: int, ablation: str = "dynamic", cfg=DEFAULT_CONFIG):
    """network/model.py:70-161, eval mode.  ``dynamic`` only needs ``dama.*`` and ``classifier.*`` keys;
    ``sfe_only`` needs ``sfe_cls.*``; ``sfe_mwt`` needs ``sfe.*``, ``mwt.*``, ``fusion_gate.*``, ``classifier.*``
    (b0 backbone: ``oracle/effnet_b0.py``)."""
    with torch.no_grad():
        if ablation == "dynamic":
            d = dama_forward(sd, "dama.", x, batch_size, cfg)
            return {"logits": classifier(sd, d["fused"]), **d}
        b, k = x.shape[:2]
        chunks = [x[:, s:min(s + batch_size, k)].flatten(0, 1) for s in range(0, k, batch_size)]   # model.py:103-105,125-127
        if ablation == "sfe_only":                 # model.py:100-118
            per_frame = torch.cat([sfe_b0_forward(sd, "sfe_cls.", c, cfg, "cls").view(b, -1, 1) for c in chunks], dim=1)
            return {"logits": per_frame.mean(dim=1), "model": "sfe_only"}
        if ablation == "sfe_mwt":                  # model.py:119-161
            dim = sd["classifier.0.weight"].shape[1]
            sfe_parts, mwt_parts = [], []
            for c in chunks:
                s_ = F.adaptive_avg_pool2d(sfe_b0_forward(sd, "sfe.", c, cfg), 1).flatten(1)         # :130-131
                sfe_parts.append(s_.view(b, -1, dim))
                mwt_parts.append(mwt_forward(sd, "mwt.", c).flatten(1).view(b, -1, dim))             # :136-138
            sfe_mean = torch.cat(sfe_parts, dim=1).mean(dim=1)                                       # :142-143
            mwt_mean = torch.cat(mwt_parts, dim=1).mean(dim=1)
            comb = torch.cat([sfe_mean, mwt_mean], dim=1)
            gate = F.relu(comb @ sd["fusion_gate.0.weight"].t() + sd["fusion_gate.0.bias"]).softmax(dim=1)   # :146-148 (ReLU, eval dropout)
            fused = sfe_mean * gate[:, 0:1] + mwt_mean * gate[:, 1:2]                                # :150-152
            return {"logits": classifier(sd, fused), "sfe": sfe_mean, "mwt": mwt_mean, "model": "sfe_mwt"}
        raise ValueError(f"Invalid ablation config: {ablation}.")
