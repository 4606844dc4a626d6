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


ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def _check_bf16(t, name, dim=None):
    _require_cuda(t, name)
    if t.dtype != torch.bfloat16 or not t.is_contiguous() or (dim is not None and t.dim() != dim):
        raise EwvitError(f"{name} must be a contiguous bf16 tensor" + (f" of rank {dim}" if dim else ""))


def _f32_or_none(t, name, n):
    if t is None:
        return None
    _require_cuda(t, name)
    if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
        raise EwvitError(f"{name} must be a contiguous fp32 tensor of {n} elements")
    return t


def linear_bf16(a, w, scale=None, shift=None, act=None, residual=None, out=None, out_dtype=torch.float32,
                splits=1, workspace=None):
    """out[M,N] = act((a[M,K] @ w[N,K]^T) * scale + shift + residual) on the tcgen05 GEMM kernel."""
    _check_bf16(a, "a", 2)
    _check_bf16(w, "w", 2)
    m, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k:
        raise EwvitError("linear_bf16: inner dimensions differ")
    scale = _f32_or_none(scale, "scale", n)
    shift = _f32_or_none(shift, "shift", n)
    if residual is not None:
        _require_cuda(residual, "residual")
        if residual.dtype != torch.float32 or residual.shape != (m, n) or residual.stride(1) != 1:
            raise EwvitError("linear_bf16: residual must be fp32 [M,N] with unit inner stride")
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    elif out.shape != (m, n) or out.stride(1) != 1 or out.dtype not in (torch.float32, torch.bfloat16):
        raise EwvitError("linear_bf16: bad out tensor")
    if splits > 1 and workspace is None:
        workspace = torch.empty((splits, m, n), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        check(load().ewvit_linear_bf16(a.data_ptr(), w.data_ptr(), m, n, k, _ptr(scale), _ptr(shift), ACT[act],
                                       _ptr(residual), residual.stride(0) if residual is not None else 0,
                                       out.data_ptr(), 1 if out.dtype == torch.float32 else 0, out.stride(0),
                                       splits, _ptr(workspace), _stream()), "ewvit_linear_bf16")
    return out


def conv3x3_bf16(x, w, n, h, wd, stride, in_padded, scale, shift, relu, y, y_coff, out_padded, force_tiled=False):
    """3x3/pad-1 conv on NHWC bf16 (see include/ewvit.h).  x, y are flat/ND contiguous buffers laid out as
    [n, h(+2), wd(+2), cin] / [n, ho(+2), wo(+2), ldc]; w is [cout, 3, 3, cin]."""
    _check_bf16(x, "x")
    _check_bf16(w, "w", 4)
    _check_bf16(y, "y")
    cout, _, _, cin = w.shape
    ldc = y.shape[-1]
    hin, win = (h + 2, wd + 2) if in_padded else (h, wd)
    if x.numel() != n * hin * win * cin:
        raise EwvitError(f"conv3x3_bf16: x has {x.numel()} elements, expected {n}x{hin}x{win}x{cin}")
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    hop, wop = (ho + 2, wo + 2) if out_padded else (ho, wo)
    if y.numel() != n * hop * wop * ldc:
        raise EwvitError(f"conv3x3_bf16: y has {y.numel()} elements, expected {n}x{hop}x{wop}x{ldc}")
    scale = _f32_or_none(scale, "scale", cout)
    shift = _f32_or_none(shift, "shift", cout)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv3x3_bf16(x.data_ptr(), w.data_ptr(), n, h, wd, cin, cout, stride, int(in_padded),
                                        _ptr(scale), _ptr(shift), int(relu), y.data_ptr(), ldc, y_coff,
                                        int(out_padded), int(force_tiled), _stream()), "ewvit_conv3x3_bf16")
    return y
