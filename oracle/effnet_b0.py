"""CPU fp32 restatement of EfficientNet-b0 ``extract_features`` (eval mode).  TEST INFRASTRUCTURE ONLY.

The reference's two ablation branches call ``EfficientNet.from_pretrained('efficientnet-b0')
.extract_features(img)`` (network/sfe.py:4,109,148; network/model.py:38-51) from the third-party
``efficientnet_pytorch`` package (lukemelas/EfficientNet-PyTorch, latest 0.7.1), which the
reference neither pins (it is not even in requirements.txt) nor vendors, and which is absent from
this image.  **PARITY UNPINNED at this boundary**: what follows restates that package's published
algorithm --

* ``utils.Conv2dStaticSamePadding``: TensorFlow 'SAME' padding computed once from the static
  input size, ``pad = max((ceil(i/s) - 1)*s + k - i, 0)``, the odd pixel going to the bottom/right;
* ``utils.get_model_params('efficientnet-b0')``: the 7-stage table below, ``se_ratio=0.25``,
  ``batch_norm_momentum=0.99`` (torch momentum 0.01), ``batch_norm_epsilon=1e-3``;
* ``model.MBConvBlock.forward``: [expand 1x1 -> BN -> swish] -> depthwise kxk -> BN -> swish ->
  SE (global mean -> reduce 1x1 + bias -> swish -> expand 1x1 + bias -> sigmoid gate) ->
  project 1x1 -> BN -> [+ input when stride 1 and in == out]; squeezed width = max(1, int(in*0.25))
  with ``in`` the block's *input* filters;
* ``model.EfficientNet.extract_features``: swish(BN(stem 3x3 s2)) -> 16 blocks -> swish(BN(head 1x1 -> 1280))

-- as a function of a state_dict with the package's key names (``_conv_stem.weight``, ``_bn0.*``,
``_blocks.N._expand_conv.weight`` ...).  It is written independently of the product's module
(``network/_effnet_b0.py``, which the golden generator runs underneath the unmodified reference
``network/model.py``) so that ``tests/test_oracle_golden.py`` is a two-implementation check.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

BN_EPS_B0 = 1e-3
# (repeats, kernel, stride, expand_ratio, input_filters, output_filters)
B0_BLOCK_ARGS = (
    (1, 3, 1, 1, 32, 16), (2, 3, 2, 6, 16, 24), (2, 5, 2, 6, 24, 40), (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112), (4, 5, 2, 6, 112, 192), (1, 3, 1, 6, 192, 320),
)


def same_pad(size: int, k: int, s: int):
    """(before, after) zero padding of TF 'SAME' for one spatial axis of static size ``size``."""
    out = math.ceil(size / s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def conv_same(x, w, b=None, stride=1, groups=1):
    k = w.shape[-1]
    ph, pw = same_pad(x.shape[-2], k, stride), same_pad(x.shape[-1], k, stride)
    if any(ph + pw):
        x = F.pad(x, (pw[0], pw[1], ph[0], ph[1]))
    return F.conv2d(x, w, b, stride=stride, groups=groups)


def _bn(x, sd: SD, p: str):
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, BN_EPS_B0)


def _swish(x):
    return x * torch.sigmoid(x)


def block_list():
    """[(kernel, stride, expand, cin, cout)] for the 16 blocks in order."""
    out = []
    for rep, k, s, e, cin, cout in B0_BLOCK_ARGS:
        for r in range(rep):
            out.append((k, s if r == 0 else 1, e, cin if r == 0 else cout, cout))
    return out


def mbconv(sd: SD, p: str, x, k, stride, expand, cin, cout):
    y = x
    if expand != 1:
        y = _swish(_bn(conv_same(y, sd[p + "_expand_conv.weight"]), sd, p + "_bn0."))
    mid = y.shape[1]
    y = _swish(_bn(conv_same(y, sd[p + "_depthwise_conv.weight"], stride=stride, groups=mid), sd, p + "_bn1."))
    g = y.mean(dim=(2, 3), keepdim=True)
    g = _swish(F.conv2d(g, sd[p + "_se_reduce.weight"], sd[p + "_se_reduce.bias"]))
    g = F.conv2d(g, sd[p + "_se_expand.weight"], sd[p + "_se_expand.bias"])
    y = torch.sigmoid(g) * y
    y = _bn(conv_same(y, sd[p + "_project_conv.weight"]), sd, p + "_bn2.")
    if stride == 1 and cin == cout:
        y = y + x
    return y


def extract_features(sd: SD, p: str, x, return_blocks=False):
    """``p`` = prefix of the EfficientNet module's keys (e.g. ``"sfe.efficient_net."``).  x [N,3,H,W] -> [N,1280,H/32,W/32]."""
    with torch.no_grad():
        y = _swish(_bn(conv_same(x, sd[p + "_conv_stem.weight"], stride=2), sd, p + "_bn0."))
        per_block = []
        for i, (k, s, e, cin, cout) in enumerate(block_list()):
            y = mbconv(sd, f"{p}_blocks.{i}.", y, k, s, e, cin, cout)
            if return_blocks:
                per_block.append(y)
        y = _swish(_bn(conv_same(y, sd[p + "_conv_head.weight"]), sd, p + "_bn1."))
    return (y, per_block) if return_blocks else y
