"""The functional oracle reproduces the UNMODIFIED reference modules (golden vectors made by
tests/golden/make_golden.py in the build container) -- SURVEY.md section 8 rows a-3 .. a-10."""
import pytest
import torch

from _weights import seeded_randn
from oracle import ewvit_oracle as O

TOL = dict(rtol=2e-4, atol=2e-4)   # fp32 CPU, different op order / threading than the reference run


@pytest.fixture(scope="module")
def frames(golden):
    x = seeded_randn((2, 3, 224, 224), golden["frames_seed"])
    assert abs(x.double().sum() - golden["frames_sum"]) < 1e-6
    return x


def test_mwt_wavelet_transform_levels(golden, dama_sd, frames):
    with torch.no_grad():
        ll, hfc, _ = O.mwt_wavelet_transform(dama_sd, "dama.mwt.", frames, (112, 112))
        g = golden["mwt_wt_l1"]
        assert torch.allclose(ll[:, :, :5, :5], g["ll_crop"], **TOL)
        assert torch.allclose(hfc[:, :8, :6, :6], g["hfc_crop"], **TOL)
        assert torch.allclose(hfc.mean(dim=(2, 3)), g["hfc_mean"], **TOL)
        _, hfc2, _ = O.mwt_wavelet_transform(dama_sd, "dama.mwt.", ll, (112, 112))
        g2 = golden["mwt_wt_l2"]
        assert torch.allclose(hfc2[:, :8, :6, :6], g2["hfc_crop"], **TOL)
        assert torch.allclose(hfc2.mean(dim=(2, 3)), g2["hfc_mean"], **TOL)


def test_hf_channels_are_colour_major(dama_sd, frames):
    """Quirk (i): hf9[:, 0:3] are the three subbands of colour 0 (mwt.py:77,85)."""
    from oracle.haar import haar_dwt2
    with torch.no_grad():
        _, _, hf9 = O.mwt_wavelet_transform(dama_sd, "dama.mwt.", frames, (112, 112))
        _, yh = haar_dwt2(frames)
    assert torch.equal(hf9[:, 0:3], yh[:, 0])       # level 1: interpolate to same size is identity
    assert torch.equal(hf9[:, 3:6], yh[:, 1])


def test_mwt_forward(golden, dama_sd, frames):
    with torch.no_grad():
        y = O.mwt_forward(dama_sd, "dama.mwt.", frames)
    assert y.shape == (2, 128, 1, 1)
    assert torch.allclose(y, golden["mwt_out"], **TOL)


def test_backbone_and_sfe(golden, dama_sd, frames):
    with torch.no_grad():
        feat = O.backbone_v2s_features(dama_sd, "dama.sfe.efficient_net.", frames)
        assert feat.shape == (2, 1280, 7, 7)
        assert torch.allclose(feat.mean(dim=(2, 3)), golden["backbone_feat_mean"], **TOL)
        y = O.sfe_head(dama_sd, "dama.sfe.", feat)
    assert y.shape == (2, 128, 1, 1)
    assert torch.allclose(y, golden["sfe_out"], **TOL)


def test_process_frame(golden, dama_sd, frames):
    with torch.no_grad():
        out = O.dama_process_frame(dama_sd, "dama.", frames)
    for k in ("fused", "space", "freq"):
        assert torch.allclose(out[k], golden["process_frame"][k], **TOL), k


def test_fusion_gate_is_centre_tap(golden, dama_sd):
    """Quirk (iv): the 3x3 fusion_gate conv on a 1x1 map only uses weight[:, :, 1, 1]."""
    g = golden["fusion_gate"]
    sd = dama_sd
    w = sd["dama.fusion_gate.0.weight"][:, :, 1, 1]
    y = g["x"].flatten(1) @ w.t() + sd["dama.fusion_gate.0.bias"]
    y = (y - sd["dama.fusion_gate.1.running_mean"]) / torch.sqrt(sd["dama.fusion_gate.1.running_var"] + 1e-5)
    y = torch.relu(y * sd["dama.fusion_gate.1.weight"] + sd["dama.fusion_gate.1.bias"])
    assert torch.allclose(y, g["y"].flatten(1), **TOL)


@pytest.mark.parametrize("case", ["detector_dynamic", "detector_config1"])
def test_detector_dynamic(golden, dama_sd, case):
    g = golden[case]
    x = seeded_randn(tuple(g["shape"]), g["seed"])
    out = O.detector_forward(dama_sd, x, g["batch_size"], "dynamic")
    for k in ("logits", "fused", "space", "freq"):
        assert out[k].shape == g[k].shape
        assert torch.allclose(out[k], g[k], **TOL), k
    assert torch.equal(out["logits"] >= 0, g["logits"] >= 0)      # identical real/fake decisions


def test_more_than_64_frames_per_chunk_raises(golden, dama_sd):
    """Quirk (ii): pos_embedding is indexed by position in the chunk; N > 64 raises (sfe.py:158-159)."""
    assert golden["n65_raises"]
    with pytest.raises(RuntimeError):
        O.vit_tokens_from_features(dama_sd, "dama.sfe.", torch.zeros(65, 1280, 7, 7))


# ---------------------------------------------------------------- ablation branches (model.py:100-161)
@pytest.fixture(scope="module")
def ablation_sd(manifest):
    from _weights import state_dict_from_manifest
    return state_dict_from_manifest(manifest, seed=0, prefixes=("sfe.", "sfe_cls.", "mwt.", "fusion_gate.", "classifier."))


def test_b0_backbone_restatement(golden, ablation_sd, frames):
    """oracle/effnet_b0.py (functional) vs the b0 module the reference ran on (two implementations of the
    published efficientnet_pytorch algorithm; upstream itself is absent: PARITY UNPINNED at that boundary)."""
    from oracle.effnet_b0 import extract_features, same_pad
    assert same_pad(224, 3, 2) == (0, 1) and same_pad(112, 3, 1) == (1, 1) and same_pad(56, 5, 2) == (1, 2)
    assert same_pad(14, 5, 1) == (2, 2) and same_pad(14, 5, 2) == (1, 2) and same_pad(7, 3, 1) == (1, 1)
    feat = extract_features(ablation_sd, "sfe.efficient_net.", frames)
    assert feat.shape == (2, 1280, 7, 7)
    assert torch.allclose(feat.mean(dim=(2, 3)), golden["b0_feat_mean"], **TOL)
    assert torch.allclose(feat[:, :16], golden["b0_feat_crop"], **TOL)


def test_sfe_b0_both_output_modes(golden, ablation_sd, frames):
    with torch.no_grad():
        y = O.sfe_b0_forward(ablation_sd, "sfe.", frames)
        c = O.sfe_b0_forward(ablation_sd, "sfe_cls.", frames, output_mode="cls")
    assert y.shape == (2, 128, 1, 1) and c.shape == (2, 1)
    assert torch.allclose(y, golden["sfe_b0_out"], **TOL)
    assert torch.allclose(c, golden["sfe_cls_out"], **TOL)


@pytest.mark.parametrize("mode", ["sfe_only", "sfe_mwt"])
@pytest.mark.parametrize("tag", ["k5", "k4"])
def test_detector_ablation_modes(golden, ablation_sd, mode, tag):
    g = golden[f"detector_{mode}_{tag}"]
    x = seeded_randn(tuple(g["shape"]), g["seed"])
    out = O.detector_forward(ablation_sd, x, g["batch_size"], mode)
    assert out["model"] == g["model"] == mode
    keys = ("logits",) if mode == "sfe_only" else ("logits", "sfe", "mwt")
    assert set(out) == set(k for k in g if k not in ("seed", "shape", "batch_size"))
    for k in keys:
        assert out[k].shape == g[k].shape
        assert torch.allclose(out[k], g[k], **TOL), k
    assert torch.equal(out["logits"] >= 0, g["logits"] >= 0)


def test_invalid_ablation_raises(ablation_sd):
    with pytest.raises(ValueError):
        O.detector_forward(ablation_sd, torch.zeros(1, 1, 3, 224, 224), 1, "nope")
