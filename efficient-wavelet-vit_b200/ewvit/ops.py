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


def _check_norm(norm, c, name):
    if norm is None or len(norm) != 2:
        raise EwvitError(f"{name}: uint8 frames need norm=(mean, std), fp32 CUDA tensors of {c} values")
    for t in norm:
        _require_cuda(t, "norm")
        if t.dtype != torch.float32 or t.numel() != c or not t.is_contiguous():
            raise EwvitError(f"{name}: norm tensors must be contiguous fp32 with {c} values")


def dwt3_haar(x: torch.Tensor, out=None, want=("ll1", "hf1", "ll2", "hf2", "ll3", "hf3"), norm=None):
    """Three chained Haar levels in one HBM pass.  x [N,C,H,W] fp32 (or uint8 with norm=(mean, std): the ToTensor +
    Normalize arithmetic happens on load), H and W multiples of 8.

    Returns a dict with the requested subbands (ll_k [N,C,H/2^k,W/2^k], hf_k [N,C,3,H/2^k,W/2^k]).
    ``out`` may carry preallocated tensors under the same keys."""
    _require_cuda(x, "x")
    u8 = x.dtype == torch.uint8
    if not (u8 or x.dtype == torch.float32) or x.dim() != 4:
        raise EwvitError("dwt3_haar: x must be fp32 or uint8 [N,C,H,W]")
    if u8:
        _check_norm(norm, x.shape[1], "dwt3_haar")
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
        if u8:
            check(load().ewvit_dwt3_haar_u8_fwd(x.data_ptr(), norm[0].data_ptr(), norm[1].data_ptr(), c, n * c, h, w, _ptr(res["ll1"]),
                                                _ptr(res["hf1"]), _ptr(res["ll2"]), _ptr(res["hf2"]), _ptr(res["ll3"]), _ptr(res["hf3"]),
                                                _stream()), "ewvit_dwt3_haar_u8_fwd")
        else:
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


def conv3x3_bf16(x, w, n, h, wd, stride, in_padded, scale, shift, relu, y, y_coff, out_padded, force_tiled=False, x_coff=0):
    """3x3/pad-1 conv on NHWC bf16 (see include/ewvit.h).  x, y are flat/ND contiguous buffers laid out as
    [n, h(+2), wd(+2), x_ldc] / [n, ho(+2), wo(+2), ldc]; w is [cout, 3, 3, cin]; the conv reads channels
    [x_coff, x_coff + cin) of x (x_ldc = x.shape[-1])."""
    _check_bf16(x, "x")
    _check_bf16(w, "w", 4)
    _check_bf16(y, "y")
    cout, _, _, cin = w.shape
    ldc = y.shape[-1]
    x_ldc = x.shape[-1] if x.dim() > 1 else cin
    hin, win = (h + 2, wd + 2) if in_padded else (h, wd)
    if x.numel() != n * hin * win * x_ldc:
        raise EwvitError(f"conv3x3_bf16: x has {x.numel()} elements, expected {n}x{hin}x{win}x{x_ldc}")
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    hop, wop = (ho + 2, wo + 2) if out_padded else (ho, wo)
    if y.numel() != n * hop * wop * ldc:
        raise EwvitError(f"conv3x3_bf16: y has {y.numel()} elements, expected {n}x{hop}x{wop}x{ldc}")
    scale = _f32_or_none(scale, "scale", cout)
    shift = _f32_or_none(shift, "shift", cout)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv3x3_bf16(x.data_ptr(), x_ldc, x_coff, w.data_ptr(), n, h, wd, cin, cout, stride, int(in_padded),
                                        _ptr(scale), _ptr(shift), int(relu), y.data_ptr(), ldc, y_coff,
                                        int(out_padded), int(force_tiled), _stream()), "ewvit_conv3x3_bf16")
    return y


def _check_f32(t, name):
    _require_cuda(t, name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise EwvitError(f"{name} must be a contiguous fp32 CUDA tensor")


def mwt_upsample3(hf1, hf2, hf3, up, h, wd):
    """hf_l [n,9,h>>(l-1),wd>>(l-1)] fp32 -> up [n,h+2,wd+2,32] bf16: channels [9l, 9l+9) = level l+1 on the h x wd grid."""
    for t, nm in ((hf1, "hf1"), (hf2, "hf2"), (hf3, "hf3")):
        _check_f32(t, nm)
    _check_bf16(up, "up")
    n = hf1.shape[0]
    if tuple(hf1.shape) != (n, 9, h, wd) or tuple(hf2.shape) != (n, 9, h // 2, wd // 2) or tuple(hf3.shape) != (n, 9, h // 4, wd // 4) \
            or up.numel() != n * (h + 2) * (wd + 2) * 32:
        raise EwvitError("mwt_upsample3: expects [n,9,h,wd], [n,9,h/2,wd/2], [n,9,h/4,wd/4] and an [n,h+2,wd+2,32] output")
    with torch.cuda.device(hf1.device):
        check(load().ewvit_mwt_upsample3_fwd(hf1.data_ptr(), hf2.data_ptr(), hf3.data_ptr(), n, h, wd, up.data_ptr(), _stream()),
              "ewvit_mwt_upsample3_fwd")
    return up


def mwt_head_conv3(up, w, scale, shift, y, h, wd):
    """up [n,h+2,wd+2,32] bf16, w [128,288] bf16, scale/shift [192] fp32 -> y [n,h+2,wd+2,192] bf16 (see include/ewvit.h)."""
    _check_bf16(up, "up")
    _check_bf16(w, "w", 2)
    _check_bf16(y, "y")
    _check_f32(scale, "scale")
    _check_f32(shift, "shift")
    n = up.numel() // ((h + 2) * (wd + 2) * 32)
    if up.numel() != n * (h + 2) * (wd + 2) * 32 or y.numel() != n * (h + 2) * (wd + 2) * 192 or tuple(w.shape) != (128, 288) \
            or scale.numel() != 192 or shift.numel() != 192:
        raise EwvitError("mwt_head_conv3: shape mismatch")
    with torch.cuda.device(up.device):
        check(load().ewvit_mwt_head_conv3_fwd(up.data_ptr(), w.data_ptr(), n, h, wd, scale.data_ptr(), shift.data_ptr(), y.data_ptr(),
                                              _stream()), "ewvit_mwt_head_conv3_fwd")
    return y


def maxpool2x2(x, y=None):
    """NHWC bf16 [n,h,w,c] -> [n,h/2,w/2,c]."""
    _check_bf16(x, "x", 4)
    n, h, w, c = x.shape
    if y is None:
        y = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_maxpool2x2_nhwc_bf16(x.data_ptr(), n, h, w, c, y.data_ptr(), _stream()),
              "ewvit_maxpool2x2_nhwc_bf16")
    return y


def gap(x, y=None):
    """NHWC bf16 [n,h,w,c] -> fp32 [n,c] mean over pixels."""
    _check_bf16(x, "x", 4)
    n, h, w, c = x.shape
    if y is None:
        y = torch.empty((n, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_gap_nhwc_bf16(x.data_ptr(), n, h * w, c, y.data_ptr(), y.stride(0), _stream()),
              "ewvit_gap_nhwc_bf16")
    return y


def vit_assemble(emb, cls, pos, pos_index, out=None):
    _check_f32(emb, "emb")
    n, d = emb.shape
    if pos_index.dtype != torch.int32 or pos_index.numel() != n or not pos_index.is_cuda:
        raise EwvitError("vit_assemble: pos_index must be an int32 CUDA tensor with one entry per frame")
    cls, pos = cls.reshape(-1), pos.reshape(-1, d)
    _check_f32(cls, "cls")
    _check_f32(pos, "pos")
    if out is None:
        out = torch.empty((n * 2, d), dtype=torch.float32, device=emb.device)
    with torch.cuda.device(emb.device):
        check(load().ewvit_vit_assemble(emb.data_ptr(), cls.data_ptr(), pos.data_ptr(), pos_index.data_ptr(), n, d,
                                        pos.shape[0], out.data_ptr(), _stream()), "ewvit_vit_assemble")
    return out


def layernorm_bf16(x, gamma, beta, eps=1e-5, out=None, rows=None, ldx=None, d=None):
    """fp32 rows -> bf16 LayerNorm (or plain cast when gamma is None). x may be a strided row view:
    pass rows/ldx/d explicitly with x the tensor whose data_ptr is the first row."""
    _require_cuda(x, "x")
    if x.dtype != torch.float32:
        raise EwvitError("layernorm_bf16: x must be fp32")
    if rows is None:
        rows, d = x.shape
        ldx = x.stride(0)
    if out is None:
        out = torch.empty((rows, d), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_layernorm_bf16(x.data_ptr(), ldx, _ptr(gamma), _ptr(beta), eps, out.data_ptr(),
                                          out.stride(0), rows, d, _stream()), "ewvit_layernorm_bf16")
    return out


def vit_attention(qkv, n, tokens, heads, dim_head, out=None):
    _check_f32(qkv, "qkv")
    if out is None:
        out = torch.empty((n * tokens, heads * dim_head), dtype=torch.bfloat16, device=qkv.device)
    with torch.cuda.device(qkv.device):
        check(load().ewvit_vit_attention(qkv.data_ptr(), n, tokens, heads, dim_head, out.data_ptr(), _stream()),
              "ewvit_vit_attention")
    return out


def dama_wpack_floats(d, depth):
    return int(load().ewvit_dama_wpack_floats(d, depth))


def dama_tail(space_in, freq_in, wpack, heads, depth, ln_eps=1e-5, out=None):
    _check_f32(space_in, "space_in")
    _check_f32(freq_in, "freq_in")
    _check_f32(wpack, "wpack")
    n, d = space_in.shape
    if wpack.numel() != dama_wpack_floats(d, depth):
        raise EwvitError("dama_tail: wpack has the wrong size")
    if out is None:
        out = tuple(torch.empty((n, d), dtype=torch.float32, device=space_in.device) for _ in range(3))
    fused, space, freq = out
    with torch.cuda.device(space_in.device):
        check(load().ewvit_dama_tail_fwd(space_in.data_ptr(), freq_in.data_ptr(), n, d, heads, depth, wpack.data_ptr(),
                                         ln_eps, fused.data_ptr(), space.data_ptr(), freq.data_ptr(), _stream()),
              "ewvit_dama_tail_fwd")
    return fused, space, freq


def video_head(fused, space, freq, videos, k, classifier=None):
    """Per-video means (+ classifier logits when classifier=(w1[hc,d], b1[hc], w2[hc], b2[1]))."""
    _check_f32(fused, "fused")
    d = fused.shape[-1]
    dev = fused.device
    mf = torch.empty((videos, d), dtype=torch.float32, device=dev)
    ms = torch.empty_like(mf) if space is not None else None
    mq = torch.empty_like(mf) if freq is not None else None
    logits = torch.empty((videos, 1), dtype=torch.float32, device=dev) if classifier is not None else None
    cw1 = cb1 = cw2 = cb2 = None
    hc = 0
    if classifier is not None:
        cw1, cb1, cw2, cb2 = classifier
        hc = cw1.shape[0]
    with torch.cuda.device(dev):
        check(load().ewvit_video_head_fwd(fused.data_ptr(), _ptr(space), _ptr(freq), videos, k, d, mf.data_ptr(),
                                          _ptr(ms), _ptr(mq), _ptr(cw1), _ptr(cb1), _ptr(cw2), _ptr(cb2), hc,
                                          _ptr(logits), _stream()), "ewvit_video_head_fwd")
    return mf, ms, mq, logits


ACT_BB = {None: 0, "none": 0, "relu": 1, "silu": 3, "silu_h": 4}      # silu_h: SiLU, weights and bias pre-halved by the caller


def conv_nhwc_bf16(x, w, ksize, stride, bias=None, act=None, residual=None, out=None):
    """Dense NHWC bf16 conv (1x1, or 3x3/pad 1) with fused bias/act/residual.  x [n,h,w,cin]; w [cout,cin] or
    [cout,9,cin_pad]; returns [n,ho,wo,cout] bf16."""
    _check_bf16(x, "x", 4)
    _check_bf16(w, "w")
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    elif out.numel() != n * ho * wo * cout or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise EwvitError("conv_nhwc_bf16: bad out tensor")
    if residual is not None:
        _check_bf16(residual, "residual")
        if residual.numel() != out.numel():
            raise EwvitError("conv_nhwc_bf16: residual must match the output")
    bias = _f32_or_none(bias, "bias", cout)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv_nhwc_bf16(x.data_ptr(), w.data_ptr(), n, h, wd, cin, cout, ksize, stride, _ptr(bias),
                                          ACT_BB[act], _ptr(residual), out.data_ptr(), _stream()), "ewvit_conv_nhwc_bf16")
    return out


def conv_nhwc_bf16_ex(x, w, ksize, stride, cin, bias=None, act=None, residual=None, out=None, in_padded=False, out_padded=False):
    """conv_nhwc_bf16 on "padded-flat" tensors: x [n,h(+2),w(+2),cin], out/residual [n,ho(+2),wo(+2),cout] (see
    ewvit_conv_nhwc_bf16_ex in include/ewvit.h; stride-2 convs with out_padded write the interior of ``out`` only)."""
    _check_bf16(x, "x", 4)
    _check_bf16(w, "w")
    n = x.shape[0]
    h, wd = (x.shape[1] - 2, x.shape[2] - 2) if in_padded else (x.shape[1], x.shape[2])
    if x.shape[3] != cin:
        raise EwvitError("conv_nhwc_bf16_ex: channel mismatch")
    cout = w.shape[0]
    ho, wo = (h - 1) // stride + 1, (wd - 1) // stride + 1
    oshape = (n, ho + 2, wo + 2, cout) if out_padded else (n, ho, wo, cout)
    if out is None:
        if out_padded and not (ksize == 1 or (stride == 1 and in_padded)):
            raise EwvitError("conv_nhwc_bf16_ex: this conv writes the interior only; pass a zero-bordered out tensor")
        out = torch.empty(oshape, dtype=torch.bfloat16, device=x.device)
    elif tuple(out.shape) != oshape or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise EwvitError("conv_nhwc_bf16_ex: bad out tensor")
    if residual is not None:
        _check_bf16(residual, "residual")
        if residual.numel() != out.numel():
            raise EwvitError("conv_nhwc_bf16_ex: residual must match the output")
    bias = _f32_or_none(bias, "bias", cout)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv_nhwc_bf16_ex(x.data_ptr(), w.data_ptr(), n, h, wd, cin, cout, ksize, stride, _ptr(bias), ACT_BB[act],
                                             _ptr(residual), out.data_ptr(), int(in_padded), int(out_padded), _stream()),
              "ewvit_conv_nhwc_bf16_ex")
    return out


def conv3x3_c24(x, w, bias, residual=False, out=None):
    """3x3/s1/p1 conv 24 -> 24 + bias + SiLU (+ x) on NHWC bf16 (ewvit_conv3x3_c24_fwd); w [24, >=216] dense tap-major."""
    _check_bf16(x, "x", 4)
    _check_bf16(w, "w", 2)
    n, h, wd, c = x.shape
    if c != 24 or w.shape[0] != 24 or w.shape[1] < 216:
        raise EwvitError("conv3x3_c24: needs 24 input/output channels and [24, >=216] weights")
    bias = _f32_or_none(bias, "bias", 24)
    if out is None:
        out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv3x3_c24_fwd(x.data_ptr(), w.data_ptr(), w.shape[1], bias.data_ptr(), n, h, wd, int(bool(residual)),
                                           out.data_ptr(), _stream()), "ewvit_conv3x3_c24_fwd")
    return out


def stem_conv(x, w, bias, out=None, out_padded=False, norm=None, same_tf=False):
    """fp32 NCHW frames [n,3,h,w] (or uint8 frames with norm=(mean, std)) -> bf16 NHWC [n,h/2,w/2,cout]: conv3x3 s2 + bias + SiLU.
    same_tf: TensorFlow 'SAME' padding (EfficientNet-b0) instead of torchvision's symmetric pad 1."""
    u8 = x.dtype == torch.uint8
    if same_tf and (u8 or out_padded):
        raise EwvitError("stem_conv: same_tf is implemented for fp32 frames and the plain output layout")
    if u8:
        _require_cuda(x, "x")
        if not x.is_contiguous() or x.dim() != 4:
            raise EwvitError("stem_conv: x must be a contiguous [n,3,h,w] tensor")
        _check_norm(norm, 3, "stem_conv")
    else:
        _check_f32(x, "x")
    _check_f32(w, "w")
    n, c, h, wd = x.shape
    cout = w.shape[0]
    if c != 3 or w.numel() != cout * 27:
        raise EwvitError("stem_conv: expects 3 input channels and [cout,3,3,3] weights")
    bias = _f32_or_none(bias, "bias", cout)
    ho, wo = (h - 1) // 2 + 1, (wd - 1) // 2 + 1
    oshape = (n, ho + 2, wo + 2, cout) if out_padded else (n, ho, wo, cout)
    if out is None:
        out = (torch.zeros if out_padded else torch.empty)(oshape, dtype=torch.bfloat16, device=x.device)
    elif tuple(out.shape) != oshape or out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise EwvitError("stem_conv: bad out tensor")
    with torch.cuda.device(x.device):
        if u8:
            check(load().ewvit_stem_conv_u8_fwd(x.data_ptr(), norm[0].data_ptr(), norm[1].data_ptr(), n, h, wd, w.data_ptr(), bias.data_ptr(),
                                                cout, out.data_ptr(), int(out_padded), _stream()), "ewvit_stem_conv_u8_fwd")
        else:
            fn = load().ewvit_stem_conv_same_fwd if same_tf else load().ewvit_stem_conv_padded_fwd if out_padded else load().ewvit_stem_conv_fwd
            check(fn(x.data_ptr(), n, h, wd, w.data_ptr(), bias.data_ptr(), cout, out.data_ptr(), _stream()), "ewvit_stem_conv_fwd")
    return out


def dwconv3x3(x, w9c, bias, stride, out=None, pooled=None):
    """Depthwise 3x3 + bias + SiLU on NHWC bf16; pooled (fp32 [n,c]) receives the spatial mean when given."""
    _check_bf16(x, "x", 4)
    _check_f32(w9c, "w9c")
    n, h, wd, c = x.shape
    bias = _f32_or_none(bias, "bias", c)
    if out is None:
        out = torch.empty((n, (h - 1) // stride + 1, (wd - 1) // stride + 1, c), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        check(load().ewvit_dwconv3x3_nhwc_bf16(x.data_ptr(), w9c.data_ptr(), bias.data_ptr(), n, h, wd, c, stride,
                                               out.data_ptr(), _ptr(pooled), _stream()), "ewvit_dwconv3x3_nhwc_bf16")
    return out


def se_gate(pooled, w1, b1, w2t, b2, out=None, bf16=False):
    """gate [n,c] = sigmoid(W2 silu(W1 mean + b1) + b2), fp32 or (bf16=True) bf16; pooled is [n,c] means or [n,parts,c]
    partial means (summed in index order)."""
    for t, nm in ((pooled, "pooled"), (w1, "w1"), (b1, "b1"), (w2t, "w2t"), (b2, "b2")):
        _check_f32(t, nm)
    n, c = pooled.shape[0], pooled.shape[-1]
    parts = pooled.shape[1] if pooled.dim() == 3 else 1
    if out is None:
        out = torch.empty((n, c), dtype=torch.bfloat16 if bf16 else torch.float32, device=pooled.device)
    with torch.cuda.device(pooled.device):
        check(load().ewvit_se_gate_fwd(pooled.data_ptr(), parts, w1.data_ptr(), b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), n, c,
                                       w1.shape[0], out.data_ptr(), int(out.dtype == torch.bfloat16), _stream()), "ewvit_se_gate_fwd")
    return out


def dwconv_out_size(size, ksize, stride, same_tf=False):
    """(output size, zero padding before) of one spatial axis: torchvision's symmetric pad k//2, or TensorFlow 'SAME'."""
    if same_tf:
        out = -(-size // stride)
        total = max((out - 1) * stride + ksize - size, 0)
        return out, total // 2
    return (size + 2 * (ksize // 2) - ksize) // stride + 1, ksize // 2


def dwconv(x, wkc, bias, ksize, stride, same_tf=False, act="silu", out=None, pooled=False):
    """Depthwise k x k (3 | 5, stride 1 | 2) + bias + SiLU on NHWC bf16 x [n,h,w,c]; wkc [k*k, c] fp32 tap-major.
    pooled=True also returns the SE squeeze as partial means [n, parts, c] fp32 (sum over parts = spatial mean)."""
    _check_bf16(x, "x", 4)
    _check_f32(wkc, "wkc")
    n, h, wd, c = x.shape
    if wkc.shape != (ksize * ksize, c):
        raise EwvitError(f"dwconv: weights must be [{ksize * ksize}, {c}] (got {tuple(wkc.shape)})")
    bias = _f32_or_none(bias, "bias", c)
    ho, pt = dwconv_out_size(h, ksize, stride, same_tf)
    wo, pl = dwconv_out_size(wd, ksize, stride, same_tf)
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    pool = None
    lib = load()
    if pooled:
        parts = lib.ewvit_dwconv_pool_parts(ho, wo, ksize, stride)
        pool = torch.empty((n, parts, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.ewvit_dwconv_nhwc_bf16(x.data_ptr(), wkc.data_ptr(), bias.data_ptr(), n, h, wd, c, ksize, stride, pt, pl, ho, wo,
                                         {None: 0, "silu": 4}[act], out.data_ptr(), _ptr(pool), _stream()), "ewvit_dwconv_nhwc_bf16")
    return (out, pool) if pooled else out


def conv1x1_gated(x, gate, w, bias=None, act=None, residual=None, out=None):
    """y = act(((x * gate[frame]) @ w^T) + bias) + residual; x [n,h,w,cin] bf16, gate [n,cin] fp32, w [cout,cin] bf16."""
    _check_bf16(x, "x", 4)
    _check_bf16(w, "w", 2)
    if gate.dtype == torch.bfloat16:
        _check_bf16(gate, "gate", 2)
    else:
        _check_f32(gate, "gate")
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    if gate.shape != (n, cin) or w.shape[1] != cin:
        raise EwvitError("conv1x1_gated: shape mismatch")
    if out is None:
        out = torch.empty((n, h, wd, cout), dtype=torch.bfloat16, device=x.device)
    if residual is not None:
        _check_bf16(residual, "residual")
        if residual.numel() != out.numel():
            raise EwvitError("conv1x1_gated: residual must match the output")
    bias = _f32_or_none(bias, "bias", cout)
    with torch.cuda.device(x.device):
        check(load().ewvit_conv1x1_gated_nhwc_bf16(x.data_ptr(), gate.data_ptr(), int(gate.dtype == torch.bfloat16), w.data_ptr(), n, h * wd, cin, cout, _ptr(bias),
                                                   ACT_BB[act], _ptr(residual), out.data_ptr(), _stream()),
              "ewvit_conv1x1_gated_nhwc_bf16")
    return out
