"""GPU parity for the tcgen05 implicit-GEMM kernel (rows a-3/a-4/a-5): C ABI vs a plain fp32 torch
reference of the same op on the same bf16-rounded operands.  Floating point: the only difference is
the fp32 accumulation order inside the tensor core, tolerance 2e-3 * max|ref| (+ bf16 output rounding
2^-8 relative when the kernel stores bf16)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from ewvit import ops
    return ops


def _randn(shape, seed, scale=1.0):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale


def _close(got, ref, bf16_out=False):
    tol = 2e-3 * float(ref.abs().max()) + 1e-6
    err = (got.float().cpu() - ref).abs()
    if bf16_out:
        err = err - ref.abs() * 2.0 ** -8
    assert float(err.max()) <= tol, f"max err {float(err.max())} > {tol}"


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (128, 128, 512), (200, 256, 512), (1024, 1536, 512),
                                   (1, 128, 64), (1024, 512, 2048), (300, 128, 128)])
def test_linear_plain(ops, m, n, k):
    a = _randn((m, k), 1).bfloat16()
    w = _randn((n, k), 2, k ** -0.5).bfloat16()
    out = ops.linear_bf16(a.cuda(), w.cuda())
    _close(out, a.float() @ w.float().t())


@pytest.mark.parametrize("act", [None, "relu", "gelu"])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_linear_epilogue(ops, act, out_dtype):
    m, n, k = 260, 256, 192
    a = _randn((m, k), 3).bfloat16()
    w = _randn((n, k), 4, k ** -0.5).bfloat16()
    scale, shift, res = _randn((n,), 5).abs() + 0.5, _randn((n,), 6), _randn((m, n), 7)
    out = ops.linear_bf16(a.cuda(), w.cuda(), scale=scale.cuda(), shift=shift.cuda(), act=act, residual=res.cuda(),
                          out_dtype=out_dtype)
    ref = (a.float() @ w.float().t()) * scale + shift + res
    ref = {None: lambda z: z, "relu": F.relu, "gelu": F.gelu}[act](ref)
    assert out.dtype == out_dtype
    _close(out, ref, bf16_out=out_dtype == torch.bfloat16)


@pytest.mark.parametrize("splits", [2, 4, 9])
def test_linear_split_k_is_deterministic(ops, splits):
    m, n, k = 512, 512, 62720 if splits == 9 else 4096
    a = _randn((m, k), 8).bfloat16().cuda()
    w = _randn((n, k), 9, k ** -0.5).bfloat16().cuda()
    bias = _randn((n,), 10).cuda()
    out1 = ops.linear_bf16(a, w, shift=bias, splits=splits)
    out2 = ops.linear_bf16(a, w, shift=bias, splits=splits)
    assert torch.equal(out1, out2)
    ref = (a.float() @ w.float().t() + bias).cpu()
    _close(out1, ref)


def test_linear_rejects_bad_shapes(ops):
    from ewvit import EwvitError
    a = torch.zeros(8, 96, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(128, 96, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(EwvitError):
        ops.linear_bf16(a, w)            # K % 64 != 0
    with pytest.raises(EwvitError):
        ops.linear_bf16(torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda"),
                        torch.zeros(100, 64, dtype=torch.bfloat16, device="cuda"))   # N % 128 != 0


def _conv_case(ops, n, h, w, cin, cout, stride, in_padded, out_padded, force_tiled, seed, ldc=None, coff=0, x_ldc=None, x_coff=0):
    x = _randn((n, cin, h, w), seed).bfloat16()
    wt = _randn((cout, cin, 3, 3), seed + 1, (9 * cin) ** -0.5).bfloat16()
    scale, shift = _randn((cout,), seed + 2).abs() + 0.5, _randn((cout,), seed + 3)
    ref = F.relu(F.conv2d(x.float(), wt.float(), stride=stride, padding=1) * scale.view(1, -1, 1, 1)
                 + shift.view(1, -1, 1, 1))
    xh = x.permute(0, 2, 3, 1).contiguous()
    if in_padded:
        xh = F.pad(xh, (0, 0, 1, 1, 1, 1))
    ho, wo = ref.shape[-2:]
    ldc = ldc or cout
    row_shift = stride == 1 and in_padded and out_padded and not force_tiled
    yshape = (n, ho + 2, wo + 2, ldc) if out_padded else (n, ho, wo, ldc)
    # the row-shift path writes its own zero border; the box path expects a pre-zeroed padded buffer
    fill = 0.0 if (out_padded and not row_shift) else 7.0
    y = torch.full(yshape, fill, dtype=torch.bfloat16, device="cuda")
    if x_ldc:       # the conv reads a channel slice of a wider tensor; the other channels hold large values that must not leak in
        wide = torch.full(xh.shape[:-1] + (x_ldc,), 100.0, dtype=torch.bfloat16)
        wide[..., x_coff:x_coff + cin] = xh
        xh = wide
    ops.conv3x3_bf16(xh.cuda().contiguous(), wt.permute(0, 2, 3, 1).contiguous().cuda(), n, h, w, stride, in_padded,
                     scale.cuda(), shift.cuda(), True, y, coff, out_padded, force_tiled, x_coff=x_coff)
    got = y.float().cpu()
    inner = got[:, 1:-1, 1:-1] if out_padded else got
    _close(inner[..., coff:coff + cout].permute(0, 3, 1, 2), ref, bf16_out=True)
    if out_padded:
        for border in (got[:, 0], got[:, -1], got[:, :, 0], got[:, :, -1]):
            assert float(border[..., coff:coff + cout].abs().max()) == 0.0, "zero border expected"
    if ldc != cout:   # channels outside [coff, coff+cout) are untouched
        mask = torch.ones(ldc, dtype=torch.bool)
        mask[coff:coff + cout] = False
        assert float((got[..., mask] - fill).abs().max()) == 0.0


@pytest.mark.parametrize("cin,cout", [(64, 128), (384, 128), (128, 128)])
def test_conv_stride1_row_shift_path(ops, cin, cout):
    _conv_case(ops, 2, 20, 24, cin, cout, 1, True, True, False, seed=20 + cin)


def test_conv_stride1_into_concat_buffer(ops):
    _conv_case(ops, 2, 12, 16, 64, 128, 1, True, True, False, seed=31, ldc=384, coff=128)


@pytest.mark.parametrize("x_coff", [0, 64, 128])
def test_conv_stride1_reads_channel_slice(ops, x_coff):
    """hf_conv.fusion reads level l of the three-level head tensor in place (x_ldc = 192, x_coff = 64 l)."""
    _conv_case(ops, 2, 12, 16, 64, 128, 1, True, True, False, seed=35 + x_coff, ldc=384, coff=128, x_ldc=192, x_coff=x_coff)


def test_conv_channel_slice_needs_the_row_shift_path(ops):
    from ewvit import EwvitError
    with pytest.raises(EwvitError):
        _conv_case(ops, 1, 12, 16, 64, 128, 2, True, False, False, seed=36, x_ldc=192, x_coff=64)


def test_conv_stride1_full_112(ops):
    _conv_case(ops, 1, 112, 112, 64, 128, 1, True, True, False, seed=33)


@pytest.mark.parametrize("in_padded", [False, True])
def test_conv_stride1_box_path_matches(ops, in_padded):
    _conv_case(ops, 2, 20, 24, 64, 128, 1, in_padded, False, True, seed=40)


@pytest.mark.parametrize("h,w,in_padded", [(28, 28, False), (112, 112, True), (30, 22, False), (9, 7, True)])
def test_conv_stride2_box_path(ops, h, w, in_padded):
    _conv_case(ops, 2, h, w, 128, 128, 2, in_padded, False, False, seed=50 + h)


def test_conv_box_path_padded_output(ops):
    _conv_case(ops, 2, 20, 24, 64, 128, 2, True, True, False, seed=60)
