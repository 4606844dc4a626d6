"""Haar analysis filter bank, CPU restatement (numpy + torch).  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: this restates ``pytorch_wavelets.DWTForward(J=1, wave='haar',
mode='zero')`` (PyPI ``pytorch-wavelets``, un-pinned by the reference,
``requirements.txt:9``; call sites ``network/mwt.py:5,20,76``).  That package is not
vendored under ``/root/reference`` and cannot be installed here, and no reference test
pins its output, so the only anchors are the published algorithm and the analytic Haar
identities in ``tests/test_oracle_haar.py``.

Published algorithm restated (``pytorch_wavelets/dwt/lowlevel.py``: ``prep_filt_afb2d``,
``afb1d``, ``AFB2D``):

* taps come from PyWavelets ``haar``: ``dec_lo = [s, s]``, ``dec_hi = [-s, s]`` with
  ``s = 1/sqrt(2)``; ``prep_filt_afb2d`` reverses them, so the cross-correlation
  kernels are ``h0 = [s, s]`` and ``h1 = [s, -s]`` (fp32, ``s = 0.70710677``);
* one level = stride-2 grouped cross-correlation along W (rows), then along H
  (columns), result reshaped to ``[N, C, 4, H/2, W/2]``; slot 0 is ``ll`` and slots
  1..3 are ``yh[..., 0..2]`` = (W-low,H-high), (W-high,H-low), (W-high,H-high);
* ``mode='zero'``: ``p = 2*(outsize-1) - N + L`` with ``outsize = (N+1)//2`` and ``L=2``
  gives ``p = 0`` for even ``N`` and ``p = 1`` for odd ``N`` -- one zero sample appended
  at the bottom / right, no other padding.

With ``a=x[2i,2j] b=x[2i,2j+1] c=x[2i+1,2j] d=x[2i+1,2j+1]`` this oracle fixes the
fp32 evaluation order (each product and each sum rounded separately, no FMA):

    lo_t = a*s + b*s      hi_t = a*s - b*s
    lo_b = c*s + d*s      hi_b = c*s - d*s
    ll = lo_t*s + lo_b*s  lh = lo_t*s - lo_b*s  hl = hi_t*s + hi_b*s  hh = hi_t*s - hi_b*s

The CUDA kernels reproduce exactly this order, so GPU-vs-oracle is bit-exact; the
oracle-vs-conv2d restatement in ``oracle/shims/pytorch_wavelets`` agrees to ~1e-7
relative (backend conv kernels may fuse multiply-adds).
"""
from __future__ import annotations

import numpy as np
import torch

HAAR_S = np.float32(0.70710677)  # fp32(1/sqrt(2)); s*s = 0.49999997 in fp32, not 0.5


def _pad_zero_mode_np(x: np.ndarray) -> np.ndarray:
    h, w = x.shape[-2:]
    ph, pw = h & 1, w & 1
    if ph or pw:
        pad = [(0, 0)] * (x.ndim - 2) + [(0, ph), (0, pw)]
        x = np.pad(x, pad)
    return x


def haar_dwt2_np(x: np.ndarray):
    """One level. x [..., H, W] float32 -> (ll [..., H2, W2], yh [..., 3, H2, W2])."""
    x = _pad_zero_mode_np(np.asarray(x, dtype=np.float32))
    s = HAAR_S
    a = x[..., 0::2, 0::2] * s
    b = x[..., 0::2, 1::2] * s
    c = x[..., 1::2, 0::2] * s
    d = x[..., 1::2, 1::2] * s
    lo_t = (a + b) * s
    hi_t = (a - b) * s
    lo_b = (c + d) * s
    hi_b = (c - d) * s
    ll = lo_t + lo_b
    lh = lo_t - lo_b
    hl = hi_t + hi_b
    hh = hi_t - hi_b
    yh = np.stack([lh, hl, hh], axis=-3)
    return ll.astype(np.float32), yh.astype(np.float32)


def haar_dwt2(x: torch.Tensor):
    """Torch twin of :func:`haar_dwt2_np` (same rounding order). x [N, C, H, W] fp32."""
    assert x.dtype == torch.float32
    h, w = x.shape[-2:]
    if (h & 1) or (w & 1):
        x = torch.nn.functional.pad(x, (0, w & 1, 0, h & 1))
    s = float(HAAR_S)
    a = x[..., 0::2, 0::2] * s
    b = x[..., 0::2, 1::2] * s
    c = x[..., 1::2, 0::2] * s
    d = x[..., 1::2, 1::2] * s
    lo_t = (a + b) * s
    hi_t = (a - b) * s
    lo_b = (c + d) * s
    hi_b = (c - d) * s
    ll = lo_t + lo_b
    yh = torch.stack([lo_t - lo_b, hi_t + hi_b, hi_t - hi_b], dim=-3)
    return ll.contiguous(), yh.contiguous()


def haar_dwt2_multilevel(x: torch.Tensor, levels: int = 3):
    """Chained levels on LL, as ``MWT.forward`` does (``network/mwt.py:104-111``).

    Returns ``[(ll_1, yh_1), ..., (ll_J, yh_J)]``.
    """
    out = []
    cur = x
    for _ in range(levels):
        ll, yh = haar_dwt2(cur)
        out.append((ll, yh))
        cur = ll
    return out
