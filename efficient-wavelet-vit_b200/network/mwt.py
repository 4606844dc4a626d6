"""``network.mwt.MWT`` -- drop-in for the reference's multi-level wavelet branch (network/mwt.py:7-119).

Same constructor, attribute names (``dwt``, ``freq_conv``, ``freq_pool``, ``hf_conv['seperate'|'fusion']``,
``multiscale_fusion``), ``wavelet_transform`` method and state_dict layout.  Eval-mode CUDA calls run the
fused native pipeline (3-level Haar kernel -> head kernel -> tcgen05 implicit-GEMM convs); training takes the
PyTorch composition below (the Haar transform itself is always the native kernel).
"""
import torch
from torch import nn
from torch.nn import functional as F

from ._native import NativeMixin


class _HaarFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from ewvit import ops
        ctx.in_shape = x.shape
        return ops.dwt_haar(x.float())

    @staticmethod
    def backward(ctx, g_ll, g_yh):
        # adjoint of the (linear) analysis bank: every output is +-s*s times one of the four inputs
        n, c, h, w = ctx.in_shape
        s2 = 0.70710677 * 0.70710677
        lh, hl, hh = g_yh[:, :, 0], g_yh[:, :, 1], g_yh[:, :, 2]
        gx = g_ll.new_zeros((n, c, 2 * g_ll.shape[-2], 2 * g_ll.shape[-1]))
        gx[..., 0::2, 0::2] = (g_ll + lh + hl + hh) * s2
        gx[..., 0::2, 1::2] = (g_ll + lh - hl - hh) * s2
        gx[..., 1::2, 0::2] = (g_ll - lh + hl - hh) * s2
        gx[..., 1::2, 1::2] = (g_ll - lh - hl + hh) * s2
        return gx[..., :h, :w]


class HaarDWT(nn.Module):
    """Stands where the reference has ``pytorch_wavelets.DWTForward(J=1, wave='haar', mode='zero')``
    (mwt.py:20): same four filter buffers in the state_dict, same ``forward(x) -> (ll, [yh])`` contract,
    computed by ``ewvit_dwt_haar_fwd``."""

    def __init__(self, J=1, wave="haar", mode="zero"):
        super().__init__()
        if J != 1 or wave != "haar" or mode != "zero":
            raise NotImplementedError("the reference uses DWTForward(J=1, wave='haar', mode='zero') only")
        s = 0.7071067811865476
        lo, hi = torch.tensor([s, s]), torch.tensor([s, -s])
        self.register_buffer("h0_col", lo.reshape(1, 1, 2, 1).clone())
        self.register_buffer("h1_col", hi.reshape(1, 1, 2, 1).clone())
        self.register_buffer("h0_row", lo.reshape(1, 1, 1, 2).clone())
        self.register_buffer("h1_row", hi.reshape(1, 1, 1, 2).clone())

    def forward(self, x):
        ll, yh = _HaarFn.apply(x)
        return ll, [yh]


def _conv_bn_relu(cin, cout, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, padding=1, stride=stride), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True))


class MWT(NativeMixin, nn.Module):
    def __init__(self, in_channels=3, dama_dim=128, levels=3):
        super().__init__()
        self.in_channels = in_channels
        self.dama_dim = dama_dim
        self.levels = levels
        self.dwt = HaarDWT(J=1, wave="haar", mode="zero")
        self.freq_conv = _conv_bn_relu(dama_dim, dama_dim, stride=2)
        self.freq_pool = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=2),
                                       nn.Conv2d(dama_dim, dama_dim, kernel_size=3, padding=1, stride=2),
                                       nn.BatchNorm2d(dama_dim), nn.ReLU(inplace=True), nn.AdaptiveAvgPool2d(1))
        self.hf_conv = nn.ModuleDict({
            "seperate": nn.ModuleList([_conv_bn_relu(in_channels, 6 * in_channels) for _ in range(3)]),
            "fusion": _conv_bn_relu(18 * in_channels, dama_dim),
        })
        self.multiscale_fusion = _conv_bn_relu(levels * dama_dim, dama_dim)

    # ---- PyTorch composition (training / hooks); mirrors mwt.py:74-90
    def wavelet_transform(self, x, target_size):
        b, c, h, w = x.shape
        ll, hf = self.dwt(x)
        hf = hf[0].reshape(b, 3 * c, hf[0].shape[-2], hf[0].shape[-1])
        if self.levels > 1:
            hf = F.interpolate(hf, size=target_size, mode="bilinear")
        groups = [self.hf_conv["seperate"][i](hf[:, i * c:(i + 1) * c]) for i in range(3)]
        return ll, self.hf_conv["fusion"](torch.cat(groups, dim=1))

    def _forward_torch(self, x):
        target = (x.shape[-2] // 2, x.shape[-1] // 2)
        cur, feats = x, []
        for _ in range(self.levels):
            cur, hf = self.wavelet_transform(cur, target)
            feats.append(hf)
        y = self.multiscale_fusion(torch.cat(feats, dim=1))
        return self.freq_pool(self.freq_conv(y))

    def _build_runner(self):
        from ewvit.engine import MwtRunner
        return MwtRunner(self.state_dict(), dim=self.dama_dim, levels=self.levels, in_channels=self.in_channels)

    def forward(self, x):
        """x [B, C, H, W] -> [B, dama_dim, 1, 1]"""
        if self._use_native(x):
            out = self._native_runner(self._build_runner).forward(x.float().contiguous())
            return out.clone().view(out.shape[0], self.dama_dim, 1, 1)
        return self._forward_torch(x)
