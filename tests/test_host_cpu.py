"""Host-side logic of the drop-in modules, no GPU: state_dict contract, seeded-init equivalence with the
reference, chunk/position rule, error behaviour, weight re-layout helpers."""
import json
import os

import pytest
import torch

from _weights import GOLDEN_DIR


@pytest.fixture(scope="module")
def model():
    from network.model import DeepfakeDetector
    torch.manual_seed(42)
    return DeepfakeDetector(in_channels=3, dama_dim=128, batch_size=8)


def test_state_dict_keys_shapes_dtypes_match_reference(model, manifest):
    """Row a-1: every one of the reference's 1743 entries, same shape and dtype (strict load both ways)."""
    sd = model.state_dict()
    assert set(sd) == set(manifest)
    for k, meta in manifest.items():
        assert list(sd[k].shape) == meta["shape"], k
        assert str(sd[k].dtype) == meta["dtype"], k
    for k in ("dama.mwt.dwt.h0_col", "dama.mwt.hf_conv.seperate.2.1.running_var", "dama.sfe.pos_embedding",
              "dama.cross_att.layers.1.3.to_kv.weight", "sfe_cls.efficient_net._blocks.15._se_expand.bias", "classifier.3.bias"):
        assert k in sd


def test_seeded_init_is_bit_identical_to_reference(model):
    """Same construction order and initialisers: torch.manual_seed(42) gives the reference's weights."""
    ref = json.load(open(os.path.join(GOLDEN_DIR, "seeded_init_sums.json")))
    sd = model.state_dict()
    bad = [k for k, s in ref.items() if abs(float(sd[k].double().sum()) - s) > 1e-9 * max(1.0, abs(s))]
    assert not bad, bad[:5]


def test_frozen_backbone_prefix(model):
    """sfe.py:115-119: the first six backbone parameters do not train."""
    flags = [p.requires_grad for _, p in model.dama.sfe.efficient_net.named_parameters()]
    assert flags[:6] == [False] * 6 and all(flags[6:])


def test_attribute_surface_used_by_reference_tools(model):
    """utils/visualize_feature_maps.py:138-168 touches these."""
    d = model.dama
    assert hasattr(d.sfe.efficient_net, "features") and hasattr(d.mwt, "dwt") and callable(d.mwt.wavelet_transform)
    assert callable(d._process_frame) and len(d.cross_att.layers) == 2 and len(d.cross_att.layers[0]) == 4
    assert d.fusion_gate[0].kernel_size == (3, 3)
    assert model.ablation_config == ["dynamic", "sfe_only", "sfe_mwt"]
    assert not hasattr(model, "ablation")          # exists only after a forward / configure_ablation (model.py:77-78)
    model.configure_ablation("sfe_mwt")
    assert model.ablation == "sfe_mwt"
    with pytest.raises(ValueError):
        model.configure_ablation("nope")
    del model.ablation


def test_eval_forward_on_cpu_fails_loudly(model):
    from ewvit import EwvitError
    model.eval()
    with pytest.raises(EwvitError, match="no CPU fallback"):
        model(torch.zeros(1, 1, 3, 224, 224), 1, "dynamic")
    with pytest.raises(EwvitError):
        model.dama.mwt(torch.zeros(1, 3, 224, 224))
    model.train()


def test_chunk_position_rule():
    """Frame (b, k) of a chunk [s, e) sits at row b*(e-s) + (k-s) of the flattened chunk (dama.py:183-186)."""
    from ewvit.engine import check_chunk_limit, chunk_pos_index
    idx = chunk_pos_index(2, 5, 2).view(2, 5)
    assert idx.tolist() == [[0, 1, 0, 1, 0], [2, 3, 2, 3, 1]]
    idx = chunk_pos_index(8, 64, 8).view(8, 64)
    assert idx[3, 17].item() == 3 * 8 + 1 and int(idx.max()) == 63
    check_chunk_limit(8, 300, 8, 64)
    with pytest.raises(RuntimeError, match=r"tensor a \(65\).*tensor b \(64\)"):
        check_chunk_limit(13, 5, 5, 64)
    check_chunk_limit(13, 4, 5, 64)      # chunks are min(batch_size, K) long


def test_fold_bn_and_tap_major_layout():
    from ewvit import engine
    torch.manual_seed(0)
    conv, bn = torch.nn.Conv2d(5, 4, 3, padding=1), torch.nn.BatchNorm2d(4)
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2)
    bn.weight.data.normal_()
    bn.bias.data.normal_()
    bn.eval()
    sd = {"c.weight": conv.weight.data, "c.bias": conv.bias.data, **{"b." + k: v for k, v in bn.state_dict().items()}}
    scale, shift = engine._fold_bn(sd, "c.", "b.")
    x = torch.randn(2, 5, 6, 6)
    raw = torch.nn.functional.conv2d(x, conv.weight, None, padding=1)
    assert torch.allclose(raw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), bn(conv(x)), atol=1e-5)
    w = engine._conv_w_tapmajor(conv.weight.data, 64)
    assert w.shape == (4, 3, 3, 64) and w.dtype == torch.bfloat16
    assert torch.equal(w[2, 1, 0, :5].float(), conv.weight.data[2, :, 1, 0].bfloat16().float())
    assert float(w[..., 5:].abs().max()) == 0.0


def test_train_mode_torch_composition_matches_oracle_in_eval_math(model, manifest):
    """The PyTorch composition kept for training is the same math as the oracle: check the DAMA fusion tail and
    the ViT head on CPU (plain sub-modules with no native kernel on their path), in eval mode."""
    from _weights import fill_module_, seeded_randn
    from oracle import ewvit_oracle as O
    fill_module_(model, seed=0)
    model.eval()
    sd = {k: v for k, v in model.state_dict().items()}
    s, f = seeded_randn((3, 1, 128), 1), seeded_randn((3, 1, 128), 2)
    with torch.no_grad():
        a, b = model.dama.cross_att(s, f)
        ra, rb = O.bidirectional_cross(sd, "dama.cross_att.", s, f, 4, 2)
        assert torch.allclose(a, ra, atol=1e-5) and torch.allclose(b, rb, atol=1e-5)
        feat = seeded_randn((2, 1280, 7, 7), 3) * 0.3
        x = O.vit_tokens_from_features(sd, "dama.sfe.", feat)
        assert torch.allclose(model.dama.sfe.transformer(x), O.vit_transformer(sd, "dama.sfe.", x), atol=1e-4)
    model.train()


def test_backbone_layout_plan_and_window_weight_packing():
    """Host logic of the native EfficientNetV2-S runner (no kernels): which ops run on padded-flat tensors, and the weight
    layout of the overlapping-window conv path (include/ewvit.h: k = dy*(nsub*64) + dx*cin + c)."""
    from torchvision.models import efficientnet_v2_s
    from ewvit import engine
    torch.manual_seed(0)
    net = efficientnet_v2_s(weights=None).eval()
    nb = engine.NativeEffNetV2(net.features, "cpu")
    kinds = [op[0] for op in nb.ops]
    assert kinds[0] == "stem" and kinds.count("dw") == 30 and len(kinds) == 140
    # stage-2 48->192 convs take the window path on padded tensors; the 24-channel stage-1 convs and everything from the
    # stride-2 conv into stage 3 on stay on plain layouts
    win = sorted(nb.win_w)
    assert win == [5, 7, 9] and all(nb.layout[i] == (True, True) for i in win)
    assert nb.layout[0] == (False, False) and nb.layout[1] == (False, False) and nb.layout[2] == (False, False)
    assert nb.layout[3] == (False, True)          # stride-2 conv writes the padded layout the next 1x1 conv keeps
    assert nb.layout[4] == (True, True) and nb.layout[10] == (True, True)
    assert nb.layout[11] == (True, False)         # stride-2 conv reads the padded interior, writes a plain tensor
    assert all(l == (False, False) for l in nb.layout[12:])
    assert not any(a != b and kinds[i] not in ("stem", "conv3") for i, (a, b) in enumerate(nb.layout))
    nb_plain = engine.NativeEffNetV2.__new__(engine.NativeEffNetV2)
    nb_plain.__dict__.update(nb.__dict__)
    assert nb_plain._plan_layouts(False) == [(False, False)] * len(nb.ops)
    # SiLU convs carry pre-halved weights (exact) and the halved activation code
    assert nb.ops[5][4] == "silu_h" and nb.ops[1][4] == "silu"          # 24->24 layers use the direct-conv kernel, not halved
    w = torch.randn(192, 48, 3, 3)
    pk = engine._w3x3_window_packed(w)
    assert pk.shape == (192, 3 * 3 * 64) and pk.dtype == torch.bfloat16
    for (oc, c, dy, dx) in ((0, 0, 0, 0), (5, 47, 2, 1), (191, 13, 1, 2)):
        assert float(pk[oc, dy * 192 + dx * 48 + c]) == float(w[oc, c, dy, dx].bfloat16())
    assert float(pk[:, 144:192].abs().max()) == 0.0 and float(pk[:, 192 + 144:384].abs().max()) == 0.0


def test_mwt_head_three_level_weight_packing(model):
    """The three-level tensor-core head's [128, 288] matrix (include/ewvit.h, ewvit_mwt_head_conv3_fwd): per tap and K step a
    [128, 16] tile; rows [0, 64) / [64, 128) of step s belong to level s / s + 1; channel 16 s + kk of a pixel is subband
    9-relative c = 16 s + kk - 9 level of that level, and row 18 g + oc holds seperate[g].weight[oc][ic][dy][dx] for c = 3 g + ic."""
    from ewvit import engine
    sd = {k[len("dama.mwt."):]: v.detach().float() for k, v in model.state_dict().items() if k.startswith("dama.mwt.")}
    run = engine.MwtRunner.__new__(engine.MwtRunner)
    try:
        engine.MwtRunner.__init__(run, sd)
    except Exception:
        pytest.skip("MwtRunner needs CUDA tensors for its remaining packs")
    w3 = run.head_w3.float().view(2, 64, 3, 3, 2, 16)          # [row half, out, dy, dx, step, kk]
    seen = torch.zeros_like(w3, dtype=torch.bool)
    for step in range(2):
        for half in range(2):
            lvl = step + half
            for g in range(3):
                wg = sd[f"hf_conv.seperate.{g}.0.weight"]
                for ic in range(3):
                    kk = 9 * lvl + 3 * g + ic - 16 * step
                    if not 0 <= kk < 16:
                        continue
                    for (oc, dy, dx) in ((0, 0, 0), (17, 2, 2), (9, 1, 0)):
                        assert float(w3[half, 18 * g + oc, dy, dx, step, kk]) == float(wg[oc, ic, dy, dx].bfloat16())
                    seen[half, 18 * g:18 * g + 18, :, :, step, kk] = True
    assert float(w3[~seen].abs().max()) == 0.0                  # everything else is zero (block-diagonal, padded)
    # every (level, subband) pair is covered exactly once over the two K steps
    cover = torch.zeros(3, 9, dtype=torch.int32)
    for step in range(2):
        for half in range(2):
            for kk in range(16):
                c = 16 * step + kk - 9 * (step + half)
                if 0 <= c < 9:
                    cover[step + half, c] += 1
    assert bool((cover == 1).all())
    assert tuple(run.head_scale192.shape) == (192,) and torch.equal(run.head_scale192[:64], run.head_scale192[128:])


def test_native_runner_cache_invalidation_rules(model):
    """The runner cache key lists (key, data_ptr, version, device) per tensor; train()/eval() transitions and
    load_state_dict drop it; `.data` writes are the documented blind spot that `invalidate_native_cache()` covers."""
    mwt = model.dama.mwt
    builds = []
    build = lambda: builds.append(1) or len(builds)
    mwt.eval()
    assert mwt._native_runner(build) == 1 and mwt._native_runner(build) == 1          # cached
    with torch.no_grad():
        mwt.freq_conv[0].weight.mul_(1.0)                                              # in-place write: version bump
    assert mwt._native_runner(build) == 2
    w = mwt.freq_conv[0].weight
    w.data = w.data.clone()                                                            # storage swap: data_ptr changes
    assert mwt._native_runner(build) == 3
    w.data.mul_(1.0)                                                                   # invisible: neither version nor pointer move
    assert mwt._native_runner(build) == 3
    mwt.invalidate_native_cache()
    assert mwt._native_runner(build) == 4
    w.data.mul_(1.0)
    model.train()
    model.eval()                                                                       # every train -> eval transition rebuilds
    assert mwt._native_runner(build) == 5
    mwt.load_state_dict(mwt.state_dict())
    assert mwt._native_runner(build) == 6
    # two tensors swapping places cannot cancel out (the old key was a SUM over tensors)
    sig = mwt._native_signature()
    assert isinstance(sig, tuple) and len(sig) == len(mwt.state_dict())
    model.train()


def test_eval_with_grad_enabled_takes_the_autograd_composition(model):
    """Eval mode + grad mode on + trainable parameters: the native outputs would carry no grad_fn, so the call is routed to
    the PyTorch composition (with a warning); under torch.no_grad() the native path is chosen (raises here: CPU tensor)."""
    import warnings
    from network import _native
    model.eval()
    cuda_like = torch.zeros(1, 3, 8, 8)

    class FakeCuda:                       # _use_native only inspects .is_cuda / .requires_grad
        is_cuda, requires_grad = True, False

    _native._WARNED.clear()
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        assert model.dama.mwt._use_native(FakeCuda()) is False
    assert any("torch.no_grad()" in str(w.message) for w in rec)
    with torch.no_grad():
        assert model.dama.mwt._use_native(FakeCuda()) is True
    for p in model.dama.mwt.parameters():
        p.requires_grad_(False)
    assert model.dama.mwt._use_native(FakeCuda()) is True      # frozen module: nothing needs a graph
    for p in model.dama.mwt.parameters():
        p.requires_grad_(True)
    del cuda_like
    model.train()


class _RecordingOps:
    """Stand-in for ``ewvit.ops`` that records the calls of a runner and returns tensors of the right shape on the CPU: the host
    orchestration (op order, layouts, channel offsets, residual wiring) is checked without a GPU."""

    def __init__(self):
        self.calls = []

    def _rec(self, name, **kw):
        self.calls.append((name, kw))

    # ---- backbone
    def stem_conv(self, frames, w, b, out=None, out_padded=False, norm=None, same_tf=False):
        self._rec("stem", padded=out_padded)
        n, _, h, wd = frames.shape
        return out if out is not None else torch.zeros(n, (h - 1) // 2 + 1, (wd - 1) // 2 + 1, w.shape[0], dtype=torch.bfloat16)

    def conv_nhwc_bf16_ex(self, x, w, k, stride, cin, bias=None, act=None, residual=None, out=None, in_padded=False, out_padded=False):
        self._rec("conv_ex", k=k, stride=stride, cin=cin, res=residual is not None, in_padded=in_padded, out_padded=out_padded)
        assert x.shape[-1] == cin
        if out is not None:
            return out
        n, h, wd = x.shape[0], x.shape[1] - 2 * in_padded, x.shape[2] - 2 * in_padded
        return torch.zeros(n, (h - 1) // stride + 1 + 2 * out_padded, (wd - 1) // stride + 1 + 2 * out_padded, w.shape[0], dtype=torch.bfloat16)

    def conv3x3_c24(self, x, w, b, residual=False):
        self._rec("c24", res=residual)
        return torch.zeros_like(x)

    def conv_nhwc_bf16(self, x, w, k, stride, bias=None, act=None, residual=None):
        self._rec("conv", k=k, stride=stride, res=residual is not None, act=act)
        if residual is not None:
            assert residual.shape[-1] == w.shape[0]
        n, h, wd, _ = x.shape
        return torch.zeros(n, (h - 1) // stride + 1, (wd - 1) // stride + 1, w.shape[0], dtype=torch.bfloat16)

    def dwconv3x3(self, x, w, b, stride, pooled=None):
        self._rec("dw", stride=stride)
        n, h, wd, c = x.shape
        assert pooled.shape == (n, c) and w.shape == (9, c)
        return torch.zeros(n, (h - 1) // stride + 1, (wd - 1) // stride + 1, c, dtype=torch.bfloat16)

    def se_gate(self, pooled, w1, b1, w2t, b2, bf16=False):
        self._rec("se", bf16=bf16)
        assert w1.shape[1] == pooled.shape[1] == w2t.shape[1]
        return torch.zeros(pooled.shape, dtype=torch.bfloat16)

    def conv1x1_gated(self, x, gate, w, bias=None, act=None, residual=None):
        self._rec("conv1g", res=residual is not None)
        assert gate.shape == (x.shape[0], x.shape[-1]) and w.shape[1] == x.shape[-1]
        n, h, wd, _ = x.shape
        return torch.zeros(n, h, wd, w.shape[0], dtype=torch.bfloat16)

    # ---- MWT
    def dwt3_haar(self, frames, out=None, want=None, norm=None):
        self._rec("dwt3", want=tuple(want))
        return out

    def mwt_upsample3(self, hf1, hf2, hf3, up, h, wd):
        self._rec("upsample3", shapes=(tuple(hf1.shape), tuple(hf2.shape), tuple(hf3.shape)), up=tuple(up.shape))
        return up

    def mwt_head_conv3(self, up, w, scale, shift, y, h, wd):
        self._rec("head_conv3", w=tuple(w.shape), scale=scale.numel(), y=tuple(y.shape))
        return y

    def conv3x3_bf16(self, x, w, n, h, wd, stride, in_padded, scale, shift, relu, y, y_coff, out_padded, force_tiled=False, x_coff=0):
        self._rec("conv3x3", x=tuple(x.shape), cin=w.shape[-1], cout=w.shape[0], stride=stride, y=tuple(y.shape), y_coff=y_coff, x_coff=x_coff,
                  in_padded=in_padded, out_padded=out_padded)
        return y

    def maxpool2x2(self, x, y=None):
        self._rec("maxpool", x=tuple(x.shape))
        return y

    def gap(self, x, y=None):
        self._rec("gap", x=tuple(x.shape))
        return torch.zeros(x.shape[0], x.shape[-1]) if y is None else y


def test_native_backbone_orchestration_with_recording_ops(monkeypatch):
    """NativeEffNetV2.forward on stubbed ops: 140 launches per forward in the torchvision order (stem, 2 direct 24->24 convs, the
    padded-flat window path of stage 2, 30 x (expand, depthwise + squeeze, SE gate, gated project)), residuals only where the block
    has a skip connection, [n, 7, 7, 1280] out."""
    from torchvision.models import efficientnet_v2_s
    from ewvit import engine
    rec = _RecordingOps()
    monkeypatch.setattr(engine, "ops", rec)
    torch.manual_seed(0)
    net = efficientnet_v2_s(weights=None).eval()
    nb = engine.NativeEffNetV2(net.features, "cpu")
    y = nb.forward(torch.zeros(2, 3, 224, 224))
    assert tuple(y.shape) == (2, 7, 7, 1280)
    names = [c[0] for c in rec.calls]
    assert len(names) == 140 and names[0] == "stem" and names[1:3] == ["c24", "c24"]
    assert names.count("dw") == names.count("se") == names.count("conv1g") == 30
    for i, nm in enumerate(names):
        if nm == "dw":
            assert names[i + 1:i + 3] == ["se", "conv1g"] and names[i - 1] == "conv"
    # skip connections: every non-first block of a stage (torchvision use_res_connect)
    want_res = sum(int(b.use_res_connect) for stage in list(net.features)[1:-1] for b in stage)
    got_res = sum(1 for nm, kw in rec.calls if kw.get("res"))
    assert got_res == want_res
    assert all(kw["bf16"] for nm, kw in rec.calls if nm == "se")


def test_mwt_runner_orchestration_with_recording_ops(model, monkeypatch):
    """MwtRunner.forward on stubbed ops: one DWT launch (HF only), ONE upsample + ONE head conv for the three levels, three fusion
    convs reading channels [64 l, 64 l + 64) of the head tensor and writing channels [128 l, 128 l + 128) of the concat buffer
    (mwt.py:113), multiscale, freq_conv (stride 2), max pool, pool conv (stride 2), GAP."""
    from ewvit import engine
    rec = _RecordingOps()
    monkeypatch.setattr(engine, "ops", rec)
    sd = {k[len("dama.mwt."):]: v.detach().float() for k, v in model.state_dict().items() if k.startswith("dama.mwt.")}
    run = engine.MwtRunner(sd)
    n = 3
    run.forward(torch.zeros(n, 3, 224, 224))
    names = [c[0] for c in rec.calls]
    assert names == ["dwt3", "upsample3", "head_conv3", "conv3x3", "conv3x3", "conv3x3", "conv3x3", "conv3x3", "maxpool", "conv3x3", "gap"]
    kw = dict(rec.calls[0][1])
    assert kw["want"] == ("hf1", "hf2", "hf3")
    up = rec.calls[1][1]
    assert up["shapes"] == ((n, 9, 112, 112), (n, 9, 56, 56), (n, 9, 28, 28)) and up["up"] == (n, 114, 114, 32)
    hc = rec.calls[2][1]
    assert hc["w"] == (128, 288) and hc["scale"] == 192 and hc["y"] == (n, 114, 114, 192)
    for lvl in range(3):
        f = rec.calls[3 + lvl][1]
        assert (f["x"], f["cin"], f["cout"], f["x_coff"], f["y_coff"], f["y"]) == ((n, 114, 114, 192), 64, 128, 64 * lvl, 128 * lvl, (n, 114, 114, 384))
        assert f["stride"] == 1 and f["in_padded"] and f["out_padded"]
    ms, fc, pc = rec.calls[6][1], rec.calls[7][1], rec.calls[9][1]
    assert (ms["cin"], ms["cout"], ms["x"], ms["y"]) == (384, 128, (n, 114, 114, 384), (n, 114, 114, 128))
    assert (fc["stride"], fc["y"], fc["in_padded"], fc["out_padded"]) == (2, (n, 56, 56, 128), True, False)
    assert (pc["stride"], pc["x"], pc["y"]) == (2, (n, 28, 28, 128), (n, 14, 14, 128))


def test_macro_batch_splitter_keeps_frame_order_and_chunk_positions(monkeypatch):
    """DamaRunner.forward_frames splits B*K frames into 512-frame passes (workspace bound); every frame keeps its (b, k) slot and its
    reference position index whatever the split (eval.py-shaped call: 8 videos x 300 frames, batch_size 8 -> 2400 frames, ragged tail)."""
    from ewvit import engine
    run = engine.DamaRunner.__new__(engine.DamaRunner)
    run.dim = 4
    run._pos_cache = {}
    run.sfe = type("S", (), {"pos": torch.zeros(64, 512)})()
    seen = []

    def fake_process(frames, pos, norm=None):
        seen.append(frames.shape[0])
        tag = frames.reshape(frames.shape[0], -1)[:, 0]                 # frame id planted in the first pixel
        f = torch.stack([tag, pos.float(), tag * 2, tag * 0 + frames.shape[0]], dim=1)
        return f, f + 1, f + 2

    monkeypatch.setattr(run, "process_frames", fake_process, raising=False)
    b, k, bs = 8, 300, 8
    x = torch.zeros(b, k, 3, 8, 8)
    x[:, :, 0, 0, 0] = torch.arange(b * k, dtype=torch.float32).view(b, k)
    fused, space, freq = run.forward_frames(x, bs)
    assert seen == [512, 512, 512, 512, 352] and fused.shape == (b * k, 4)
    assert torch.equal(fused[:, 0], torch.arange(b * k, dtype=torch.float32))      # order preserved
    assert torch.equal(space, fused + 1) and torch.equal(freq, fused + 2)
    want_pos = engine.chunk_pos_index(b, k, bs).float()
    assert torch.equal(fused[:, 1], want_pos)
    # the last reference chunk holds 300 - 37*8 = 4 frames per video: positions b*4 + j
    assert want_pos.view(b, k)[3, 296:].tolist() == [12.0, 13.0, 14.0, 15.0]
    with pytest.raises(RuntimeError):                                             # 9 videos x 8 frames per chunk = 72 > 64 rows
        run.forward_frames(torch.zeros(9, 16, 3, 8, 8), 8)


def test_sfe_head_orchestration_with_recording_ops(model, monkeypatch):
    """SfeRunner.head on stubbed ops: split-K patch embedding, token assembly, then per layer LN -> qkv -> 2-token attention ->
    out-projection(+residual) -> LN -> ff1(GELU) -> ff2(+residual), final cast of token 1 and the feat_map Linear + ReLU (sfe.py:153-173)."""
    from ewvit import engine
    from network._native import load_architecture_config
    calls = []

    class Ops:
        def linear_bf16(self, a, w, scale=None, shift=None, act=None, residual=None, out=None, out_dtype=None, splits=1, workspace=None):
            calls.append(("linear", tuple(a.shape), tuple(w.shape), act, residual is not None, splits))
            return out if out is not None else torch.zeros(a.shape[0], w.shape[0])

        def vit_assemble(self, emb, cls, pos, pos_index, out=None):
            calls.append(("assemble", tuple(emb.shape), tuple(out.shape)))
            return out

        def layernorm_bf16(self, x, g, b, eps=1e-5, out=None, rows=None, ldx=None, d=None):
            calls.append(("ln", g is not None, rows))
            return out

        def vit_attention(self, qkv, n, tokens, heads, dim_head, out=None):
            calls.append(("attn", n, tokens, heads, dim_head))
            return out

    monkeypatch.setattr(engine, "ops", Ops())
    cfg = load_architecture_config()
    sd = {k[len("dama.sfe."):]: v.detach() for k, v in model.state_dict().items()
          if k.startswith("dama.sfe.") and not k.startswith("dama.sfe.efficient_net.")}
    run = engine.SfeRunner(sd, cfg, backbone=None)
    n = 5
    run.head(torch.zeros(n, 62720, dtype=torch.bfloat16), torch.arange(n, dtype=torch.int32))
    kinds = [c[0] for c in calls]
    layer = ["ln", "linear", "attn", "linear", "ln", "linear", "linear"]
    assert kinds == ["linear", "assemble"] + layer * 2 + ["ln", "linear"]
    assert calls[0][1:3] == ((n, 62720), (512, 62720)) and calls[0][5] > 1                  # split-K patch_to_embedding
    qkv, out_p, ff1, ff2 = calls[3], calls[5], calls[7], calls[8]
    assert qkv[1:3] == ((2 * n, 512), (1536, 512)) and not qkv[4]
    assert out_p[2] == (512, 512) and out_p[4] and ff1[2] == (2048, 512) and ff1[3] == "gelu" and ff2[2] == (512, 2048) and ff2[4]
    assert calls[4] == ("attn", n, 2, 8, 64)
    assert calls[-2] == ("ln", False, n) and calls[-1][2] == (128, 512) and calls[-1][3] == "relu"


def test_head3_packing_reproduces_the_per_colour_convs_in_the_kernels_algebra(model):
    """CPU emulation of what ewvit_mwt_head_conv3_fwd computes from `pack_head3_weights`: a pixel row holds channel 9 l + c; per tap the
    K step 0 (channels 0..15) multiplies the [level 0 | level 1] tile into accumulator columns [0, 128) and K step 1 (channels 16..31)
    the [level 1 | level 2] tile into columns [64, 192).  The result must equal the reference's three per-colour Conv2d(3 -> 18) of
    every level (mwt.py:84-86, before BatchNorm/ReLU)."""
    import torch.nn.functional as F
    from ewvit import engine
    sd = {k[len("dama.mwt."):]: v.detach().float() for k, v in model.state_dict().items() if k.startswith("dama.mwt.")}
    run = engine.MwtRunner(sd)
    w3 = run.head_w3.float().view(128, 9, 2, 16)                          # [row, tap, step, kk]
    g = torch.Generator().manual_seed(5)
    h = w = 10
    hf = [torch.randn(1, 9, h, w, generator=g).bfloat16().float() for _ in range(3)]      # upsampled subbands of the three levels
    up = torch.zeros(1, h + 2, w + 2, 32)
    for lvl in range(3):
        up[:, 1:-1, 1:-1, 9 * lvl:9 * lvl + 9] = hf[lvl].permute(0, 2, 3, 1)
    acc = torch.zeros(h, w, 192)
    for dy in range(3):
        for dx in range(3):
            win = up[0, dy:dy + h, dx:dx + w]                              # [h, w, 32] = the tap's shifted window
            tap = dy * 3 + dx
            acc[..., 0:128] += win[..., 0:16] @ w3[:, tap, 0].t()
            acc[..., 64:192] += win[..., 16:32] @ w3[:, tap, 1].t()
    for lvl in range(3):
        ref = torch.cat([F.conv2d(hf[lvl][:, 3 * i:3 * i + 3], sd[f"hf_conv.seperate.{i}.0.weight"].bfloat16().float(), padding=1)
                         for i in range(3)], dim=1)[0].permute(1, 2, 0)     # [h, w, 54], bias/BN live in the epilogue's scale/shift
        got = acc[..., 64 * lvl:64 * lvl + 54]
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
        assert float(acc[..., 64 * lvl + 54:64 * lvl + 64].abs().max()) == 0.0
