"""Tensor-level wrappers over the C ABI: validate, allocate outputs with torch, pass raw pointers.

PyTorch is plumbing here (device memory, streams); every arithmetic op below runs in libewvit.so.
"""
import torch

from ._lib import EwvitError, check, load


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise EwvitError(f"{name} must be a CUDA tensor: libewvit.so has no CPU path")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def dwt_haar(x: torch.Tensor):
    """One Haar level, zero-mode boundary.  x [N,C,H,W] fp32 -> (ll [N,C,H2,W2], yh [N,C,3,H2,W2]).

    Same contract as ``pytorch_wavelets.DWTForward(J=1,'haar','zero')(x)`` -> ``(ll, [yh])``
    at reference network/mwt.py:76."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 4:
        raise EwvitError("dwt_haar: x must be fp32 [N,C,H,W]")
    x = x.contiguous()
    n, c, h, w = x.shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    ll = torch.empty((n, c, h2, w2), dtype=torch.float32, device=x.device)
    yh = torch.empty((n, c, 3, h2, w2), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_dwt_haar_fwd(x.data_ptr(), n * c, h, w, ll.data_ptr(), yh.data_ptr(), _stream()),
              "ewvit_dwt_haar_fwd")
    return ll, yh


def dwt3_haar(x: torch.Tensor, out=None, want=("ll1", "hf1", "ll2", "hf2", "ll3", "hf3")):
    """Three chained Haar levels in one HBM pass.  x [N,C,H,W] fp32, H and W multiples of 8.

    Returns a dict with the requested subbands (ll_k [N,C,H/2^k,W/2^k], hf_k [N,C,3,H/2^k,W/2^k]).
    ``out`` may carry preallocated tensors under the same keys."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 4:
        raise EwvitError("dwt3_haar: x must be fp32 [N,C,H,W]")
    x = x.contiguous()
    n, c, h, w = x.shape
    res = {}
    for lvl in (1, 2, 3):
        hh, ww = h >> lvl, w >> lvl
        for kind, shape in (("ll", (n, c, hh, ww)), ("hf", (n, c, 3, hh, ww))):
            key = f"{kind}{lvl}"
            if key not in want:
                res[key] = None
            elif out is not None and key in out:
                t = out[key]
                if t.shape != shape or t.dtype != torch.float32 or not t.is_contiguous() or t.device != x.device:
                    raise EwvitError(f"dwt3_haar: bad preallocated tensor for {key}")
                res[key] = t
            else:
                res[key] = torch.empty(shape, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_dwt3_haar_fwd(x.data_ptr(), n * c, h, w, _ptr(res["ll1"]), _ptr(res["hf1"]),
                                         _ptr(res["ll2"]), _ptr(res["hf2"]), _ptr(res["ll3"]), _ptr(res["hf3"]),
                                         _stream()), "ewvit_dwt3_haar_fwd")
    return {k: v for k, v in res.items() if v is not None}
