"""EfficientNet-b0 feature extractor with ``efficientnet_pytorch``-compatible state_dict keys.

The reference's two ablation branches build ``EfficientNet.from_pretrained('efficientnet-b0')``
from the third-party ``efficientnet_pytorch`` package (network/sfe.py:4,109,148), which is
neither pinned nor vendored.  When that package is importable ``network/sfe.py`` uses it; when
it is not (this image), this module supplies the same architecture under the same parameter
names (``_conv_stem``, ``_bn0``, ``_blocks.N._expand_conv`` ... ``_conv_head``, ``_bn1``,
``_fc``) so that checkpoints keep loading with ``strict=True``.  Pretrained ImageNet weights
cannot be downloaded here; initialisation is random.
"""
import math

import torch
from torch import nn
from torch.nn import functional as F

# (repeats, kernel, stride, expand, in, out) -- the b0 stage table; se_ratio 0.25 throughout
_B0_STAGES = (
    (1, 3, 1, 1, 32, 16),
    (2, 3, 2, 6, 16, 24),
    (2, 5, 2, 6, 24, 40),
    (3, 3, 2, 6, 40, 80),
    (3, 5, 1, 6, 80, 112),
    (4, 5, 2, 6, 112, 192),
    (1, 3, 1, 6, 192, 320),
)
_BN_MOMENTUM = 0.01
_BN_EPS = 1e-3
_DROP_CONNECT = 0.2


class _SamePadConv2d(nn.Conv2d):
    """TF-style 'same' padding fixed for a known input size (extra pixel goes bottom/right)."""

    def __init__(self, cin, cout, kernel_size, image_size, stride=1, groups=1, bias=False):
        super().__init__(cin, cout, kernel_size, stride=stride, groups=groups, bias=bias)
        k, s = kernel_size, stride
        out = math.ceil(image_size / s)
        total = max((out - 1) * s + k - image_size, 0)
        self._pad = (total // 2, total - total // 2, total // 2, total - total // 2)
        self.out_size = out

    def forward(self, x):
        if any(self._pad):
            x = F.pad(x, self._pad)
        return F.conv2d(x, self.weight, self.bias, self.stride, 0, self.dilation, self.groups)


def _bn(c):
    return nn.BatchNorm2d(c, momentum=_BN_MOMENTUM, eps=_BN_EPS)


class _MBConv(nn.Module):
    def __init__(self, cin, cout, kernel, stride, expand, image_size):
        super().__init__()
        mid = cin * expand
        self._has_expand = expand != 1
        if self._has_expand:
            self._expand_conv = _SamePadConv2d(cin, mid, 1, image_size)
            self._bn0 = _bn(mid)
        self._depthwise_conv = _SamePadConv2d(mid, mid, kernel, image_size, stride=stride, groups=mid)
        self._bn1 = _bn(mid)
        size = self._depthwise_conv.out_size
        squeezed = max(1, int(cin * 0.25))
        self._se_reduce = _SamePadConv2d(mid, squeezed, 1, 1, bias=True)
        self._se_expand = _SamePadConv2d(squeezed, mid, 1, 1, bias=True)
        self._project_conv = _SamePadConv2d(mid, cout, 1, size)
        self._bn2 = _bn(cout)
        self._skip = stride == 1 and cin == cout
        self.out_size = size

    def forward(self, x, drop_rate=0.0):
        y = x
        if self._has_expand:
            y = F.silu(self._bn0(self._expand_conv(y)))
        y = F.silu(self._bn1(self._depthwise_conv(y)))
        g = F.adaptive_avg_pool2d(y, 1)
        g = self._se_expand(F.silu(self._se_reduce(g)))
        y = torch.sigmoid(g) * y
        y = self._bn2(self._project_conv(y))
        if self._skip:
            if self.training and drop_rate > 0:
                keep = 1.0 - drop_rate
                mask = torch.floor(keep + torch.rand(y.shape[0], 1, 1, 1, dtype=y.dtype, device=y.device))
                y = y / keep * mask
            y = y + x
        return y


class EfficientNet(nn.Module):
    """b0 only.  ``extract_features(img[N,3,224,224]) -> [N,1280,7,7]`` (sfe.py:148)."""

    def __init__(self, image_size=224, num_classes=1000):
        super().__init__()
        self._conv_stem = _SamePadConv2d(3, 32, 3, image_size, stride=2)
        self._bn0 = _bn(32)
        size = self._conv_stem.out_size
        blocks = []
        for repeats, k, s, e, cin, cout in _B0_STAGES:
            for r in range(repeats):
                blk = _MBConv(cin if r == 0 else cout, cout, k, s if r == 0 else 1, e, size)
                size = blk.out_size
                blocks.append(blk)
        self._blocks = nn.ModuleList(blocks)
        self._conv_head = _SamePadConv2d(320, 1280, 1, size)
        self._bn1 = _bn(1280)
        self._avg_pooling = nn.AdaptiveAvgPool2d(1)
        self._dropout = nn.Dropout(0.2)
        self._fc = nn.Linear(1280, num_classes)

    @classmethod
    def from_pretrained(cls, model_name, **kwargs):
        if model_name != "efficientnet-b0":
            raise NotImplementedError("only efficientnet-b0 is on the reference's path (sfe.py:109)")
        net = cls(**kwargs)        # this stand-in has no weight source: random init (network/sfe.py warns about it)
        net._ewvit_random_init = True
        return net

    @classmethod
    def from_name(cls, model_name, **kwargs):
        if model_name != "efficientnet-b0":
            raise NotImplementedError("only efficientnet-b0 is on the reference's path (sfe.py:109)")
        return cls(**kwargs)

    def extract_features(self, x):
        x = F.silu(self._bn0(self._conv_stem(x)))
        n = len(self._blocks)
        for i, blk in enumerate(self._blocks):
            x = blk(x, _DROP_CONNECT * i / n)
        return F.silu(self._bn1(self._conv_head(x)))

    def forward(self, x):
        x = self._avg_pooling(self.extract_features(x)).flatten(1)
        return self._fc(self._dropout(x))
