"""GPU parity of the native forward path against the CPU fp32 oracle on the same seeded inputs and the same
key-addressed weights (rows a-3 .. a-10), plus the committed golden outputs of the unmodified reference.

Floating point, bf16 tensor-core operands with fp32 accumulation: the stated tolerance is
``|got - ref| <= tol * max|ref|`` per tensor (measured errors are printed by every check):
* TOL = 1.5e-2: MWT branch tensors (measured 2e-3..7e-3), the ViT head on identical features (4e-3) and EVERY output of
  the full-size runs (512 and 2400 frames, per-video means over 64 / 300 frames: 3e-3..6e-3);
* TOL_FRAME = 3e-2: outputs that sit behind the ~170-layer bf16 EfficientNet chain and are NOT averaged over a full
  video -- per-frame `_process_frame` outputs, per-video means over <= 8 frames and the logits computed from them
  (measured 7e-3..2.6e-2; the backbone feature map itself is within 2.9e-2 of torchvision fp32, cuDNN's own bf16 path
  within 3.9e-2); TOL_B0 = 4e-2 for the same kind of outputs behind the EfficientNet-b0 chain of the ablation branches
  (16 SE-gated MBConv blocks with bf16 gates: per-frame features measured 3.1e-2, 4-5-frame logits of `sfe_only` 3.1e-2);
* 1e-4 for the fp32-only DAMA tail; and identical
real/fake decisions (sign of the logit) wherever |logit_ref| exceeds the tolerance."""
import pytest
import torch

from _weights import seeded_randn
from oracle import ewvit_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1.5e-2
TOL_FRAME = 3e-2
TOL_B0 = 4e-2


def rel_err(got, ref):
    return float((got.detach().float().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-12))


def check(name, got, ref, tol=TOL):
    e = rel_err(got, ref)
    print(f"[parity] {name}: max|err|/max|ref| = {e:.3e} (tol {tol:g})")
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert e <= tol, f"{name}: {e} > {tol}"


@pytest.fixture(scope="module")
def sd_cuda(dama_sd):
    return {k: v.cuda() for k, v in dama_sd.items()}


@pytest.fixture(scope="module")
def frames(golden):
    return seeded_randn((2, 3, 224, 224), golden["frames_seed"])


@pytest.fixture(scope="module")
def detector(manifest):
    """The drop-in module tree with the key-addressed weights, on the GPU, eval mode."""
    from _weights import fill_module_
    from network.model import DeepfakeDetector
    m = DeepfakeDetector(3, 128, batch_size=8)
    fill_module_(m, seed=0)
    return m.cuda().eval()


def test_mwt_head_three_levels_in_one_launch(dama_sd, sd_cuda, frames):
    """What MwtRunner runs: one upsample launch (channel 9 l + c of a 64-byte pixel row) + one block-diagonal conv whose tile serves
    the three levels from one window fetch ([level][64] channels per pixel), against the per-colour convs of the oracle (mwt.py:77-86)."""
    from ewvit import engine, ops
    run = engine.MwtRunner(engine._sub(sd_cuda, "dama.mwt."))
    with torch.no_grad():
        _, inter = O.mwt_forward(dama_sd, "dama.mwt.", frames, return_intermediates=True)
    out = ops.dwt3_haar(frames.cuda(), want=("hf1", "hf2", "hf3"))
    hfs = [out[f"hf{l}"].view(2, 9, 224 >> l, 224 >> l) for l in (1, 2, 3)]
    up = torch.zeros((2, 114, 114, 32), dtype=torch.bfloat16, device="cuda")
    y = torch.full((2, 114, 114, 192), 7.0, dtype=torch.bfloat16, device="cuda")          # borders must come out as zeros
    ops.mwt_upsample3(*hfs, up, 112, 112)
    upc = up.float().cpu()
    ops.mwt_head_conv3(up, run.head_w3, run.head_scale192, run.head_shift192, y, 112, 112)
    got = y.float().cpu()
    for lvl in range(3):
        hf9 = inter[f"hf9_{lvl}"]
        check(f"mwt_upsample3 level {lvl + 1}", upc[:, 1:-1, 1:-1, 9 * lvl:9 * lvl + 9].permute(0, 3, 1, 2), hf9, 1e-2)
        parts = [O._conv_bn_relu(hf9[:, 3 * i:3 * i + 3], dama_sd, f"dama.mwt.hf_conv.seperate.{i}.0.",
                                 f"dama.mwt.hf_conv.seperate.{i}.1.") for i in range(3)]
        check(f"mwt_head_conv3 level {lvl + 1}", got[:, 1:-1, 1:-1, 64 * lvl:64 * lvl + 54].permute(0, 3, 1, 2), torch.cat(parts, dim=1), 2e-2)
        assert float(got[:, 1:-1, 1:-1, 64 * lvl + 54:64 * lvl + 64].abs().max()) == 0.0
    assert float(upc[..., 27:].abs().max()) == 0.0 and float(upc[:, 0].abs().max()) == 0.0 and float(upc[:, :, -1].abs().max()) == 0.0
    assert float(got[:, 0].abs().max()) == 0.0 and float(got[:, :, -1].abs().max()) == 0.0 and float(got[:, -1].abs().max()) == 0.0

def test_mwt_forward(dama_sd, sd_cuda, frames, golden):
    from ewvit import engine
    run = engine.MwtRunner(engine._sub(sd_cuda, "dama.mwt."))
    got = run.forward(frames.cuda())
    with torch.no_grad():
        ref, inter = O.mwt_forward(dama_sd, "dama.mwt.", frames, return_intermediates=True)
    ws = run._workspace(2, 224, 224)
    for lvl in range(3):
        check(f"hf_conv.fusion level {lvl + 1}", ws["cat"][:, 1:-1, 1:-1, lvl * 128:(lvl + 1) * 128].permute(0, 3, 1, 2),
              inter[f"hfc_{lvl}"])
    check("multiscale_fusion", ws["ms"][:, 1:-1, 1:-1].permute(0, 3, 1, 2), inter["multiscale"])
    check("freq_conv", ws["fc"].permute(0, 3, 1, 2), inter["freq_conv"])
    check("MWT.forward", got, ref.flatten(1))
    check("MWT.forward vs reference golden", got, golden["mwt_out"].flatten(1))
    # the padded buffers keep their zero border (the next conv relies on it)
    assert float(ws["cat"][:, 0].abs().max()) == 0.0 and float(ws["ms"][:, :, 0].abs().max()) == 0.0


def test_mwt_module_matches_runner_and_ragged_batch(detector, dama_sd):
    """`MWT.forward` through the drop-in module; odd batch size; same result twice (workspace reuse)."""
    x = seeded_randn((3, 3, 224, 224), 77)
    with torch.no_grad():
        got1 = detector.dama.mwt(x.cuda())
        got2 = detector.dama.mwt(x.cuda())
        ref = O.mwt_forward(dama_sd, "dama.mwt.", x)
    assert got1.shape == (3, 128, 1, 1) and torch.equal(got1, got2)
    check("MWT module", got1, ref)


def test_native_b0_backbone(detector, manifest, frames, golden):
    """EfficientNet-b0 `extract_features` on the native kernels (TF-SAME stem, 3x3 / 5x5 depthwise, SE-gated project convs) vs the
    functional fp32 oracle (oracle/effnet_b0.py) and the golden features of the module the reference ran on."""
    from _weights import state_dict_from_manifest
    from ewvit import engine
    from oracle.effnet_b0 import extract_features
    sd = state_dict_from_manifest(manifest, seed=0, prefixes=("sfe.efficient_net.",))
    bb = engine.NativeEffNetB0(detector.sfe.efficient_net, "cuda")
    with torch.no_grad():
        f = bb.forward(frames.cuda())
        ref = extract_features(sd, "sfe.efficient_net.", frames)
    assert f.shape == (2, 7, 7, 1280)
    check("EfficientNet-b0 features (native)", f.permute(0, 3, 1, 2), ref, TOL_B0)
    check("EfficientNet-b0 feature means vs reference golden", f.float().mean(dim=(1, 2)), golden["b0_feat_mean"], TOL_B0)


def test_sfe_head_on_identical_features(dama_sd, sd_cuda, frames):
    """patch_to_embedding (split-K) + 2-token ViT + feat_map on the SAME bf16-rounded backbone features."""
    from ewvit import engine
    with torch.no_grad():
        feat = O.backbone_v2s_features(dama_sd, "dama.sfe.efficient_net.", frames).bfloat16()
        ref = O.sfe_head(dama_sd, "dama.sfe.", feat.float())
    sub = {k: v for k, v in engine._sub(sd_cuda, "dama.sfe.").items() if not k.startswith("efficient_net.")}
    run = engine.SfeRunner(sub, O.DEFAULT_CONFIG, backbone=None)
    nhwc = feat.permute(0, 2, 3, 1).reshape(2, -1).contiguous().cuda()
    got = run.head(nhwc, torch.arange(2, dtype=torch.int32, device="cuda"))
    check("SFE head", got, ref.flatten(1))


def test_sfe_head_position_index(dama_sd, sd_cuda):
    """Quirk (ii): frame f gets pos_embedding[pos_index[f]]; permuting the index changes the output accordingly."""
    from ewvit import engine
    feat = (seeded_randn((4, 1280, 7, 7), 5) * 0.3).bfloat16()
    sub = {k: v for k, v in engine._sub(sd_cuda, "dama.sfe.").items() if not k.startswith("efficient_net.")}
    run = engine.SfeRunner(sub, O.DEFAULT_CONFIG, backbone=None)
    nhwc = feat.permute(0, 2, 3, 1).reshape(4, -1).contiguous().cuda()
    idx = torch.tensor([2, 0, 3, 1], dtype=torch.int32)
    got = run.head(nhwc, idx.cuda()).cpu()
    with torch.no_grad():
        # oracle: put each frame at the chunk position its index names
        order = torch.argsort(idx)
        ref_sorted = O.sfe_head(dama_sd, "dama.sfe.", feat.float()[order]).flatten(1)
    check("SFE head with permuted positions", got[order], ref_sorted)


def test_dama_tail_fp32(dama_sd, sd_cuda):
    from ewvit import engine, ops
    space, freq = seeded_randn((7, 128, 1, 1), 61).abs(), seeded_randn((7, 128, 1, 1), 62).abs()
    with torch.no_grad():
        ref = O.dama_fuse(dama_sd, "dama.", space, freq)
    wpack = engine.pack_dama_weights(engine._sub(sd_cuda, "dama."), 128, 2)
    fused, s, f = ops.dama_tail(space.flatten(1).cuda().contiguous(), freq.flatten(1).cuda().contiguous(), wpack, 4, 2)
    check("DAMA fused", fused, ref["fused"], 1e-4)
    check("DAMA space", s, ref["space"], 1e-4)
    check("DAMA freq", f, ref["freq"], 1e-4)


def test_process_frame(detector, dama_sd, frames, golden):
    with torch.no_grad():
        got = detector.dama._process_frame(frames.cuda())
        ref = O.dama_process_frame(dama_sd, "dama.", frames)
    for k in ("fused", "space", "freq"):
        check(f"_process_frame[{k}]", got[k], ref[k], TOL_FRAME)
        check(f"_process_frame[{k}] vs reference golden", got[k], golden["process_frame"][k], TOL_FRAME)


@pytest.mark.parametrize("case", ["detector_dynamic", "detector_config1"])
def test_detector_dynamic_matches_reference_golden(detector, dama_sd, golden, case):
    """Full `model(x, batch_size, 'dynamic')`: ragged last chunk (K=5, bs=2) and BASELINE config 1."""
    g = golden[case]
    x = seeded_randn(tuple(g["shape"]), g["seed"])
    with torch.no_grad():
        out = detector(x.cuda(), g["batch_size"], "dynamic")
    ref = O.detector_forward(dama_sd, x, g["batch_size"], "dynamic")
    assert sorted(out) == ["freq", "fused", "logits", "space"]
    for k in ("fused", "space", "freq", "logits"):
        tol_k = TOL_FRAME                                    # means over 5 / 8 frames
        check(f"{case}[{k}] vs oracle", out[k], ref[k], tol_k)
        check(f"{case}[{k}] vs reference golden", out[k], g[k], tol_k)
    tol = TOL_FRAME * float(g["logits"].abs().max())
    decided = g["logits"].abs() > tol
    assert torch.equal((out["logits"].cpu() >= 0)[decided], (g["logits"] >= 0)[decided])


def test_detector_many_videos_one_pass(detector, dama_sd):
    """8 videos x 6 frames, batch_size 3: 48 frames in one native pass == oracle's 2 serial chunks of 24."""
    x = seeded_randn((8, 6, 3, 224, 224), 91)
    with torch.no_grad():
        out = detector(x.cuda(), 3, "dynamic")
    ref = O.detector_forward(dama_sd, x, 3, "dynamic")
    check("8x6 videos [fused]", out["fused"], ref["fused"], TOL_FRAME)
    check("8x6 videos [logits]", out["logits"], ref["logits"], TOL_FRAME)


def test_full_size_config3_matches_oracle_and_is_deterministic(detector, dama_sd):
    """BASELINE configs[2] at full size: 8 videos x 64 frames, batch_size 8 (reference: 8 serial chunks of 64 frames; here one
    native pass of 512 frames).  Logits / fused features against the fp32 CPU oracle within the bf16 tolerance, identical
    real/fake decisions wherever the reference logit is farther from 0 than the tolerance, and bit-identical results
    across two runs (no atomics, fixed reduction orders)."""
    x = seeded_randn((8, 64, 3, 224, 224), 97)
    xc = x.cuda()
    with torch.no_grad():
        a = detector(xc, 8, "dynamic")
        b = detector(xc, 8, "dynamic")
    for k in ("logits", "fused", "space", "freq"):
        assert torch.equal(a[k], b[k]), f"{k}: two runs differ"
    ref = O.detector_forward(dama_sd, x, 8, "dynamic")
    for k in ("fused", "space", "freq", "logits"):
        check(f"512 frames [{k}]", a[k], ref[k])
    tol = TOL * float(ref["logits"].abs().max())
    decided = ref["logits"].abs() > tol
    assert torch.equal((a["logits"].cpu() >= 0)[decided], (ref["logits"] >= 0)[decided])


def test_chunk_limit_raises_like_reference(detector):
    """Quirk (ii): B * batch_size > 64 frames per chunk raises RuntimeError (sfe.py:158-159)."""
    with pytest.raises(RuntimeError, match="must match the size"):
        with torch.no_grad():
            detector(torch.zeros(13, 5, 3, 224, 224, device="cuda"), 5, "dynamic")


def test_load_state_dict_invalidates_native_cache(detector, dama_sd, frames):
    from _weights import fill_module_
    with torch.no_grad():
        a = detector.dama.mwt(frames.cuda()).clone()
        fill_module_(detector, seed=1)
        b = detector.dama.mwt(frames.cuda()).clone()
        fill_module_(detector, seed=0)
        c = detector.dama.mwt(frames.cuda()).clone()
    assert not torch.equal(a, b) and torch.equal(a, c)


@pytest.fixture(scope="module")
def ablation_sd(manifest):
    from _weights import state_dict_from_manifest
    return state_dict_from_manifest(manifest, seed=0, prefixes=("sfe.", "sfe_cls.", "mwt.", "fusion_gate.", "classifier."))


@pytest.mark.parametrize("tag", ["k5", "k4"])
@pytest.mark.parametrize("mode", ["sfe_only", "sfe_mwt"])
def test_detector_ablation_modes_match_reference_golden(detector, ablation_sd, golden, mode, tag):
    """`model(x, batch_size, 'sfe_only' | 'sfe_mwt')` (model.py:100-161; b0 branches): same dict keys, values within the
    bf16 tolerance of the outputs of the unmodified reference (golden) and of the oracle, identical decisions."""
    g = golden[f"detector_{mode}_{tag}"]
    x = seeded_randn(tuple(g["shape"]), g["seed"])
    with torch.no_grad():
        out = detector(x.cuda(), g["batch_size"], mode)
    ref = O.detector_forward(ablation_sd, x, g["batch_size"], mode)
    assert out["model"] == mode and sorted(out) == sorted(k for k in g if k not in ("seed", "shape", "batch_size"))
    assert detector.ablation == mode
    for k in (("logits",) if mode == "sfe_only" else ("sfe", "mwt", "logits")):
        tol_k = TOL if k == "mwt" or (k == "logits" and mode == "sfe_mwt") else TOL_B0       # b0 chain, means over 4-5 frames
        check(f"{mode}/{tag}[{k}] vs oracle", out[k], ref[k], tol_k)
        check(f"{mode}/{tag}[{k}] vs reference golden", out[k], g[k], tol_k)
    tol = TOL_B0 * float(g["logits"].abs().max())
    decided = g["logits"].abs() > tol
    assert torch.equal((out["logits"].cpu() >= 0)[decided], (g["logits"] >= 0)[decided])


def test_sfe_b0_module_both_output_modes(detector, ablation_sd, frames, golden):
    """Standalone `EfficientViT` on the EfficientNet-b0 backbone (sfe.py:109,148), feature-map and cls modes."""
    with torch.no_grad():
        y = detector.sfe(frames.cuda())
        c = detector.sfe_cls(frames.cuda())
    assert y.shape == (2, 128, 1, 1) and c.shape == (2, 1)
    check("sfe (b0, feature_map) vs reference golden", y, golden["sfe_b0_out"], TOL_B0)      # per frame, 80-layer bf16 b0 chain
    check("sfe_cls (b0, cls) vs reference golden", c, golden["sfe_cls_out"], TOL_B0)


def test_full_shape_config4_eval_scoring_call(detector, dama_sd):
    """BASELINE configs[3] shape of ONE eval call (eval.py:135-194): 8 videos x 300 frames, batch_size 8 -> the reference runs
    37 chunks of 64 frames + a ragged chunk of 32; here 2400 frames go through the 512-frame macro-batch splitter (4 x 512 +
    352, macro-batch boundaries fall inside videos and inside chunks).  Per-video logits / features vs the fp32 CPU oracle."""
    x = seeded_randn((8, 300, 3, 224, 224), 101)
    with torch.no_grad():
        out = detector(x.cuda(), 8, "dynamic")
        torch.cuda.synchronize()
    ref = O.detector_forward(dama_sd, x, 8, "dynamic")
    for k in ("fused", "space", "freq", "logits"):
        check(f"8x300 frames [{k}]", out[k], ref[k])
    tol = TOL * float(ref["logits"].abs().max())
    decided = ref["logits"].abs() > tol
    assert torch.equal((out["logits"].cpu() >= 0)[decided], (ref["logits"] >= 0)[decided])


def test_forward_uint8_equals_forward_on_normalised_frames(detector):
    """forward_uint8 (normalisation fused into the DWT and stem kernels) is bit-identical to forward on host-normalised frames"""
    g = torch.Generator().manual_seed(11)
    u = torch.randint(0, 256, (2, 5, 3, 224, 224), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 1, 3, 1, 1)
    xf = u.float().div(255).sub(mean).div(std)
    with torch.no_grad():
        ref = detector(xf.cuda(), 2, "dynamic")
        got = detector.forward_uint8(u.cuda(), 2)
    for k in ("logits", "fused", "space", "freq"):
        assert torch.equal(got[k], ref[k]), k
    from ewvit import EwvitError
    with pytest.raises(EwvitError):
        detector.forward_uint8(xf.cuda(), 2)


def test_other_dama_dims():
    """dama_dim = 256 runs natively (two 128-wide column tiles per MWT conv; looser tolerance: the committed BatchNorm
    calibration scalars are for the reference's default dim = 128); a dim the kernels do not tile raises loudly."""
    from _weights import fill_module_
    from ewvit import EwvitError
    from network.model import DeepfakeDetector
    torch.manual_seed(0)
    m = DeepfakeDetector(3, 256, batch_size=2)
    fill_module_(m, seed=0)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items() if k.startswith("dama.") or k.startswith("classifier.")}
    m = m.cuda().eval()
    x = seeded_randn((1, 2, 3, 224, 224), 5)
    with torch.no_grad():
        out = m(x.cuda(), 2, "dynamic")
    ref = O.detector_forward(sd, x, 2, "dynamic")
    for k in ("logits", "fused", "space", "freq"):
        check(f"dim 256 [{k}]", out[k], ref[k], 6e-2)
    m64 = DeepfakeDetector(3, 64, batch_size=2).cuda().eval()
    with pytest.raises(EwvitError, match="multiple of 128"):
        with torch.no_grad():
            m64(x.cuda(), 2, "dynamic")


def test_training_micro_step_matches_reference_golden(golden):
    """Row f-2 pinned to the reference: ONE training micro-step (train mode, BatchNorm batch statistics per chunk of B*batch_size
    frames, dropout / stochastic depth probabilities 0, focal + orthogonality loss, backward) of the drop-in modules on the GPU
    against the same step of the UNMODIFIED reference on the CPU (tests/golden/make_golden.py): outputs, both loss terms, a
    spread of parameter gradients (classifier, DAMA, ViT, MWT, backbone), updated BatchNorm buffers, and which parameters get no
    gradient.  fp32 on both sides (TF32 off): tolerance 2e-3 of the tensor's max (different conv algorithms / reduction orders
    through ~170 BatchNorm layers with batch statistics of 4 frames)."""
    from _weights import fill_module_
    from ewvit.training import binary_focal_loss, orthogonal_loss
    from network.model import DeepfakeDetector
    g = golden["train_step"]
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(0)
        m = DeepfakeDetector(3, 128, batch_size=2)
        fill_module_(m, seed=0)
        m = m.cuda().train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout) or type(mod).__name__ == "StochasticDepth":
                mod.p = 0.0
        x = seeded_randn(tuple(g["shape"]), g["seed"]).cuda()
        out = m(x, g["batch_size"], "dynamic")
        cls = binary_focal_loss(out["logits"], g["labels"].cuda().view(-1, 1))
        orth = orthogonal_loss(out["space"], out["freq"])
        (cls + orth).backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for k in ("logits", "fused", "space", "freq"):
        check(f"train step [{k}]", out[k], g[k], 2e-3)
    cls, orth = cls.detach(), orth.detach()
    assert abs(float(cls) - g["cls_loss"]) <= 2e-3 * abs(g["cls_loss"]) + 1e-7, (float(cls), g["cls_loss"])
    assert abs(float(orth) - g["orth_loss"]) <= 1e-2 * abs(g["orth_loss"]) + 1e-7, (float(orth), g["orth_loss"])
    named = dict(m.named_parameters())
    for k, ref in g["grads"].items():
        if g["grad_norms"][k] < 1e-4:          # conv biases in front of a train-mode BatchNorm: the gradient is rounding noise
            continue
        got = named[k].grad
        got = got if got.numel() <= 4096 else got.flatten()[:4096]
        check(f"train step grad[{k}]", got, ref, 1e-2)
    for k in g["no_grad"]:
        assert named[k].grad is None, k
    sd = m.state_dict()
    for k, ref in g["buffers"].items():
        if ref.dtype == torch.int64:
            assert int(sd[k]) == int(ref), k
        else:
            check(f"train step buffer[{k}]", sd[k], ref, 2e-3)


def test_training_step_focal_loss_accumulation():
    """BASELINE configs[4], one rank: two accumulated micro-steps (forward + backward through the PyTorch composition with the
    native Haar kernel and its adjoint, focal + orthogonality loss) and one Adam step.  Gradients reach the dynamic path
    only (model.mwt / model.sfe / model.sfe_cls / model.fusion_gate stay untouched), the step moves the
    weights and the result stays finite; afterwards the eval-mode native path serves the updated weights."""
    from _weights import fill_module_
    from ewvit.training import train_step
    from network.model import DeepfakeDetector
    torch.manual_seed(0)
    m = DeepfakeDetector(3, 128, batch_size=2)
    fill_module_(m, seed=0)
    m = m.cuda().train()
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4)
    before = m.dama.mwt.multiscale_fusion[0].weight.detach().clone()
    micro = [(seeded_randn((1, 2, 3, 224, 224), 60 + i).cuda(), torch.tensor([i], device="cuda")) for i in range(2)]
    loss = train_step(m, micro, opt, batch_size=2, epoch=5, max_epochs=10)
    assert loss == loss and abs(loss) < 1e4
    assert m.dama.mwt.multiscale_fusion[0].weight.grad is not None and m.classifier[0].weight.grad is not None
    assert m.mwt.multiscale_fusion[0].weight.grad is None and m.fusion_gate[0].weight.grad is None
    assert not torch.equal(before, m.dama.mwt.multiscale_fusion[0].weight.detach())
    m.eval()
    with torch.no_grad():
        out = m(micro[0][0], 2, "dynamic")
    assert torch.isfinite(out["logits"]).all()
