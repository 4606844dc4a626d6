"""GPU parity for the native EfficientNetV2-S feature extractor (row a-6 / f-1): each kernel against a plain fp32
torch reference on the same bf16-rounded operands, then the whole extractor against torchvision's own fp32 module
(the oracle's backbone) with the bf16 tolerance of tests/test_model_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from _weights import seeded_randn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from ewvit import ops
    return ops


def _close(got, ref, tol=2e-3, bf16_out=True):
    err = (got.float().cpu() - ref).abs()
    if bf16_out:
        err = err - ref.abs() * 2.0 ** -8
    lim = tol * float(ref.abs().max()) + 1e-6
    assert float(err.max()) <= lim, f"max err {float(err.max())} > {lim}"


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("cin,cout,act,res", [(24, 48, None, False), (96, 48, None, True), (160, 960, "silu", False),
                                              (1536, 256, None, True), (256, 1280, "silu", False), (64, 256, "silu", False)])
def test_conv1x1(ops, cin, cout, act, res):
    n, h, w = 3, 14, 10
    x = seeded_randn((n, cin, h, w), 1).bfloat16()
    wt = (seeded_randn((cout, cin, 1, 1), 2) * cin ** -0.5).bfloat16()
    b = seeded_randn((cout,), 3)
    r = seeded_randn((n, cout, h, w), 4).bfloat16() if res else None
    ref = F.conv2d(x.float(), wt.float(), b)
    if act == "silu":
        ref = F.silu(ref)
    if res:
        ref = ref + r.float()
    got = ops.conv_nhwc_bf16(_nhwc(x).cuda(), wt.flatten(1).contiguous().cuda(), 1, 1, bias=b.cuda(), act=act,
                             residual=_nhwc(r).cuda() if res else None)
    _close(got.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("cin,cout,stride,res,hw", [(24, 24, 1, True, (112, 112)), (24, 96, 2, False, (112, 112)),
                                                   (48, 192, 1, False, (56, 56)), (64, 256, 1, False, (28, 28)),
                                                   (48, 192, 2, False, (20, 12))])
def test_conv3x3_small_channels(ops, cin, cout, stride, res, hw):
    from ewvit import engine
    n, (h, w) = 2, hw
    x = seeded_randn((n, cin, h, w), 5).bfloat16()
    wt = (seeded_randn((cout, cin, 3, 3), 6) * (9 * cin) ** -0.5).bfloat16()
    b = seeded_randn((cout,), 7)
    ref = F.silu(F.conv2d(x.float(), wt.float(), b, stride=stride, padding=1))
    if res:
        ref = ref + x.float()
    wk = engine._w3x3_tapmajor_padded(wt.float()).cuda()
    got = ops.conv_nhwc_bf16(_nhwc(x).cuda(), wk, 3, stride, bias=b.cuda(), act="silu", residual=_nhwc(x).cuda() if res else None)
    _close(got.permute(0, 3, 1, 2), ref)


@pytest.mark.parametrize("res,hw", [(True, (112, 112)), (False, (20, 12)), (True, (17, 33))])
def test_conv3x3_c24_direct(ops, res, hw):
    """warp-level-MMA direct conv of the 24 -> 24 stage-1 blocks == F.conv2d + SiLU (+ x)"""
    from ewvit import engine
    n, (h, w) = 3, hw
    x = seeded_randn((n, 24, h, w), 41).bfloat16()
    wt = (seeded_randn((24, 24, 3, 3), 42) * (9 * 24) ** -0.5).bfloat16()
    b = seeded_randn((24,), 43)
    ref = F.silu(F.conv2d(x.float(), wt.float(), b, padding=1))
    if res:
        ref = ref + x.float()
    got = ops.conv3x3_c24(_nhwc(x).cuda(), engine._w3x3_tapmajor_padded(wt.float()).cuda(), b.cuda(), residual=res)
    _close(got.permute(0, 3, 1, 2), ref)


def test_stem(ops):
    x = seeded_randn((3, 3, 224, 224), 8)
    wt = seeded_randn((24, 3, 3, 3), 9) * 27 ** -0.5
    b = seeded_randn((24,), 10)
    ref = F.silu(F.conv2d(x, wt, b, stride=2, padding=1))
    got = ops.stem_conv(x.cuda(), wt.cuda().contiguous(), b.cuda())
    assert got.shape == (3, 112, 112, 24)
    _close(got.permute(0, 3, 1, 2), ref, tol=1e-4)


def _pad_nhwc(x_nchw):
    """NCHW -> padded-flat NHWC [n, h+2, w+2, c] with a zero border."""
    return F.pad(_nhwc(x_nchw), (0, 0, 1, 1, 1, 1)).contiguous()


def _check_padded(got, ref, tol=2e-3):
    """interior == ref, border == 0"""
    _close(got[:, 1:-1, 1:-1].permute(0, 3, 1, 2), ref, tol=tol)
    g = got.float()
    assert float(g[:, 0].abs().max()) == 0 and float(g[:, -1].abs().max()) == 0
    assert float(g[:, :, 0].abs().max()) == 0 and float(g[:, :, -1].abs().max()) == 0


@pytest.mark.parametrize("cin,cout,res,hw", [(24, 24, True, (112, 112)), (48, 192, False, (56, 56)), (16, 64, False, (20, 12)),
                                            (40, 264, False, (9, 30))])
def test_conv3x3_window_path(ops, cin, cout, res, hw):
    """overlapping-window TMA path (padded-flat in and out, cin < 64, stride 1) == F.conv2d"""
    from ewvit import engine
    n, (h, w) = 2, hw
    x = seeded_randn((n, cin, h, w), 31).bfloat16()
    wt = (seeded_randn((cout, cin, 3, 3), 32) * (9 * cin) ** -0.5).bfloat16()
    b = seeded_randn((cout,), 33)
    ref = F.silu(F.conv2d(x.float(), wt.float(), b, padding=1))
    if res:
        ref = ref + x.float()
    xp = _pad_nhwc(x).cuda()
    got = ops.conv_nhwc_bf16_ex(xp, engine._w3x3_window_packed(wt.float()).cuda(), 3, 1, cin, bias=b.cuda(), act="silu",
                                residual=xp if res else None, in_padded=True, out_padded=True)
    assert got.shape == (n, h + 2, w + 2, cout)
    _check_padded(got.cpu(), ref)


def test_conv1x1_padded_layout(ops):
    n, cin, cout, h, w = 3, 96, 48, 14, 10
    x = seeded_randn((n, cin, h, w), 34).bfloat16()
    wt = (seeded_randn((cout, cin, 1, 1), 35) * cin ** -0.5).bfloat16()
    b = seeded_randn((cout,), 36) + 1.0       # a non-zero bias would leak into the border without the masking
    r = seeded_randn((n, cout, h, w), 37).bfloat16()
    ref = F.conv2d(x.float(), wt.float(), b) + r.float()
    got = ops.conv_nhwc_bf16_ex(_pad_nhwc(x).cuda(), wt.flatten(1).contiguous().cuda(), 1, 1, cin, bias=b.cuda(),
                                residual=_pad_nhwc(r).cuda(), in_padded=True, out_padded=True)
    _check_padded(got.cpu(), ref)


@pytest.mark.parametrize("cin,cout,out_padded,hw", [(24, 96, True, (112, 112)), (48, 192, False, (56, 56)), (24, 96, True, (20, 12))])
def test_conv3x3_stride2_from_padded_input(ops, cin, cout, out_padded, hw):
    from ewvit import engine
    n, (h, w) = 2, hw
    x = seeded_randn((n, cin, h, w), 38).bfloat16()
    wt = (seeded_randn((cout, cin, 3, 3), 39) * (9 * cin) ** -0.5).bfloat16()
    b = seeded_randn((cout,), 40)
    ref = F.silu(F.conv2d(x.float(), wt.float(), b, stride=2, padding=1))
    ho, wo = ref.shape[2:]
    out = torch.zeros((n, ho + 2, wo + 2, cout), dtype=torch.bfloat16, device="cuda") if out_padded else None
    got = ops.conv_nhwc_bf16_ex(_pad_nhwc(x).cuda(), engine._w3x3_tapmajor_padded(wt.float()).cuda(), 3, 2, cin, bias=b.cuda(),
                                act="silu", out=out, in_padded=True, out_padded=out_padded)
    if out_padded:
        _check_padded(got.cpu(), ref)
    else:
        _close(got.permute(0, 3, 1, 2), ref)


def test_stem_tf_same_padding(ops):
    """EfficientNet-b0 stem: 3 -> 32, 3x3 stride 2 with TensorFlow 'SAME' padding (zero row/column at the bottom/right only)."""
    x = seeded_randn((2, 3, 224, 224), 8)
    wt = seeded_randn((32, 3, 3, 3), 9) * 27 ** -0.5
    b = seeded_randn((32,), 10)
    ref = F.silu(F.conv2d(F.pad(x, (0, 1, 0, 1)), wt, b, stride=2))
    got = ops.stem_conv(x.cuda(), wt.cuda().contiguous(), b.cuda(), same_tf=True)
    assert got.shape == (2, 112, 112, 32)
    _close(got.permute(0, 3, 1, 2), ref, tol=1e-3)


def test_stem_padded(ops):
    x = seeded_randn((2, 3, 224, 224), 8)
    wt = seeded_randn((24, 3, 3, 3), 9) * 27 ** -0.5
    b = seeded_randn((24,), 10)
    ref = F.silu(F.conv2d(x, wt, b, stride=2, padding=1))
    got = ops.stem_conv(x.cuda(), wt.cuda().contiguous(), b.cuda(), out_padded=True)
    assert got.shape == (2, 114, 114, 24)
    _check_padded(got.cpu(), ref, tol=1e-4)


@pytest.mark.parametrize("c,stride,hw", [(256, 2, (28, 28)), (960, 1, (14, 14)), (1536, 1, (7, 7)), (960, 2, (14, 14)), (64, 1, (5, 9))])
def test_depthwise_and_squeeze(ops, c, stride, hw):
    n, (h, w) = 3, hw
    x = seeded_randn((n, c, h, w), 11).bfloat16()
    wt = seeded_randn((c, 1, 3, 3), 12) * (1 / 3)
    b = seeded_randn((c,), 13) * 0.1
    ref = F.silu(F.conv2d(x.float(), wt, b, stride=stride, padding=1, groups=c))
    pooled = torch.empty((n, c), device="cuda")
    got = ops.dwconv3x3(_nhwc(x).cuda(), wt.reshape(c, 9).t().contiguous().cuda(), b.cuda(), stride, pooled=pooled)
    _close(got.permute(0, 3, 1, 2), ref, tol=1e-4)
    _close(pooled, ref.mean(dim=(2, 3)), tol=5e-3, bf16_out=False)


@pytest.mark.parametrize("c,k,stride,hw,same_tf", [
    (256, 3, 2, (28, 28), False), (960, 3, 1, (14, 14), False), (1536, 3, 1, (7, 7), False), (960, 3, 2, (14, 14), False),
    (64, 3, 1, (5, 9), False), (512, 3, 1, (14, 14), False), (768, 3, 1, (14, 14), False),
    # EfficientNet-b0 shapes: TensorFlow 'SAME' padding, 5x5 kernels, channel counts that are not multiples of 64
    (32, 3, 1, (112, 112), True), (96, 3, 2, (112, 112), True), (144, 5, 2, (56, 56), True), (240, 5, 1, (28, 28), True),
    (240, 3, 2, (28, 28), True), (480, 5, 1, (14, 14), True), (672, 5, 2, (14, 14), True), (1152, 5, 1, (7, 7), True),
    (1152, 3, 1, (7, 7), True), (40, 5, 2, (9, 11), True)])
def test_depthwise_general(ops, c, k, stride, hw, same_tf):
    """ewvit_dwconv_nhwc_bf16: k x k depthwise + bias + SiLU with torchvision (symmetric) or TF-SAME padding, plus the SE
    squeeze as partial means -- vs fp32 torch on the same bf16 input (fp32 accumulation, tanh.approx SiLU: 1e-3)."""
    n, (h, w) = 3, hw
    x = seeded_randn((n, c, h, w), 11).bfloat16()
    wt = seeded_randn((c, 1, k, k), 12) * (1.0 / k)
    b = seeded_randn((c,), 13) * 0.1
    if same_tf:
        (ho, pt), (wo, pl) = ops.dwconv_out_size(h, k, stride, True), ops.dwconv_out_size(w, k, stride, True)
        pb, pr = max((ho - 1) * stride + k - h, 0) - pt, max((wo - 1) * stride + k - w, 0) - pl
        ref = F.silu(F.conv2d(F.pad(x.float(), (pl, pr, pt, pb)), wt, b, stride=stride, groups=c))
    else:
        ref = F.silu(F.conv2d(x.float(), wt, b, stride=stride, padding=k // 2, groups=c))
    got, pooled = ops.dwconv(_nhwc(x).cuda(), wt.reshape(c, k * k).t().contiguous().cuda(), b.cuda(), k, stride, same_tf=same_tf,
                             pooled=True)
    assert got.shape == (n, ref.shape[2], ref.shape[3], c)
    _close(got.permute(0, 3, 1, 2), ref, tol=1e-3)
    assert pooled.shape[0] == n and pooled.shape[2] == c
    _close(pooled.sum(dim=1), ref.mean(dim=(2, 3)), tol=5e-3, bf16_out=False)
    # the gate kernel consumes the partial means directly
    sq = 8
    w1, b1 = seeded_randn((sq, c), 15) * c ** -0.5, seeded_randn((sq,), 16)
    w2, b2 = seeded_randn((c, sq), 17) * sq ** -0.5, seeded_randn((c,), 18)
    if c % 8 == 0:
        gate = ops.se_gate(pooled, w1.cuda(), b1.cuda(), w2.t().contiguous().cuda(), b2.cuda())
        gref = torch.sigmoid(F.silu(pooled.sum(dim=1).cpu() @ w1.t() + b1) @ w2.t() + b2)
        _close(gate, gref, tol=1e-3, bf16_out=False)


@pytest.mark.parametrize("c,cout,hw,res", [(960, 160, (14, 14), True), (1536, 256, (7, 7), True), (256, 128, (14, 14), False),
                                           (768, 160, (14, 14), False)])
def test_gated_project_conv(ops, c, cout, hw, res):
    """SE gate fused into the project conv's A path == scale pass followed by the plain 1x1 conv."""
    n, (h, w) = 5, hw
    x = seeded_randn((n, c, h, w), 21).bfloat16()
    gate = torch.sigmoid(seeded_randn((n, c), 22))
    wt = (seeded_randn((cout, c, 1, 1), 23) * c ** -0.5).bfloat16()
    b = seeded_randn((cout,), 24)
    r = seeded_randn((n, cout, h, w), 25).bfloat16() if res else None
    xs = (x.float() * gate.view(n, c, 1, 1)).bfloat16().float()      # the kernel rounds the scaled operand to bf16
    ref = F.conv2d(xs, wt.float(), b)
    if res:
        ref = ref + r.float()
    got = ops.conv1x1_gated(_nhwc(x).cuda(), gate.cuda().contiguous(), wt.flatten(1).contiguous().cuda(), bias=b.cuda(),
                            residual=_nhwc(r).cuda() if res else None)
    _close(got.permute(0, 3, 1, 2), ref)
    # bf16 gates (what the engine uses): same contract with the gate rounded to bf16 first
    gb = gate.bfloat16()
    xs = (x.float() * gb.float().view(n, c, 1, 1)).bfloat16().float()
    ref = F.conv2d(xs, wt.float(), b)
    if res:
        ref = ref + r.float()
    got = ops.conv1x1_gated(_nhwc(x).cuda(), gb.cuda().contiguous(), wt.flatten(1).contiguous().cuda(), bias=b.cuda(),
                            residual=_nhwc(r).cuda() if res else None)
    _close(got.permute(0, 3, 1, 2), ref)


def test_se_gate(ops):
    n, c, sq = 6, 960, 40
    pooled = seeded_randn((n, c), 26)
    w1, b1 = seeded_randn((sq, c), 27) * c ** -0.5, seeded_randn((sq,), 28)
    w2, b2 = seeded_randn((c, sq), 29) * sq ** -0.5, seeded_randn((c,), 30)
    ref = torch.sigmoid(F.silu(pooled @ w1.t() + b1) @ w2.t() + b2)
    got = ops.se_gate(pooled.cuda(), w1.cuda(), b1.cuda(), w2.t().contiguous().cuda(), b2.cuda())
    _close(got, ref, tol=1e-3, bf16_out=False)
    got = ops.se_gate(pooled.cuda(), w1.cuda(), b1.cuda(), w2.t().contiguous().cuda(), b2.cuda(), bf16=True)
    assert got.dtype == torch.bfloat16
    _close(got, ref, tol=1e-3, bf16_out=True)


def test_native_backbone_matches_torchvision(dama_sd, golden):
    from ewvit import engine
    from oracle import ewvit_oracle as O
    from torchvision.models import efficientnet_v2_s
    frames = seeded_randn((2, 3, 224, 224), golden["frames_seed"])
    net = efficientnet_v2_s(weights=None)
    net.classifier = torch.nn.Identity()
    p = "dama.sfe.efficient_net."
    net.load_state_dict({k[len(p):]: v for k, v in dama_sd.items() if k.startswith(p)})
    net.eval()
    nb = engine.NativeEffNetV2(net.features, "cuda")
    got = nb.forward(frames.cuda())
    assert got.shape == (2, 7, 7, 1280) and got.dtype == torch.bfloat16
    with torch.no_grad():
        ref = O.backbone_v2s_features(dama_sd, p, frames)
    e = float((got.float().cpu().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max())
    print(f"[parity] native EfficientNetV2-S features: max|err|/max|ref| = {e:.3e} (tol 6e-2)")
    assert e <= 6e-2
    assert torch.allclose(got.float().cpu().permute(0, 3, 1, 2).mean(dim=(2, 3)), golden["backbone_feat_mean"], atol=2e-2, rtol=5e-2)
