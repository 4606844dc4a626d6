#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Makes the reference importable exactly as SURVEY.md section 8c describes: two shim
packages for the missing third-party modules (``oracle/shims``), CWD = /root/reference so
that ``config/architecture.yaml`` resolves, and the ImageNet weight download disabled.
Writes

* ``state_dict_manifest.json`` -- every key / shape / dtype of ``DeepfakeDetector(3,128)``
  (the drop-in's state_dict contract; ``sfe.*`` / ``sfe_cls.*`` backbone keys come from the
  b0 stand-in, not from the real ``efficientnet_pytorch``),
* ``ewvit_golden.pt`` -- outputs of the reference modules on seeded inputs with the
  key-addressed weights of ``tests/_weights.py``.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.normpath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("EWVIT_REFERENCE", "/root/reference")

sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, os.path.join(REPO, "oracle", "shims"))
sys.path.insert(0, REF)
os.chdir(REF)

import torch  # noqa: E402
import torchvision.models as tvm  # noqa: E402

from _weights import fill_module_, seeded_randn  # noqa: E402

# network/sfe.py:111-112 asks for ImageNet weights (a download); force random init instead.
_orig_v2s = tvm.efficientnet_v2_s
import network.sfe as ref_sfe  # noqa: E402

ref_sfe.efficientnet_v2_s = lambda weights=None, **kw: _orig_v2s(weights=None, **kw)

from network.model import DeepfakeDetector  # noqa: E402
from pytorch_wavelets import DWTForward  # noqa: E402

torch.manual_seed(42)
torch.set_num_threads(os.cpu_count() or 1)

model = DeepfakeDetector(in_channels=3, dama_dim=128, batch_size=8)
manifest = {k: {"shape": list(v.shape), "dtype": str(v.dtype)} for k, v in model.state_dict().items()}
with open(os.path.join(HERE, "state_dict_manifest.json"), "w") as f:
    json.dump(manifest, f, indent=0, sort_keys=True)
print("manifest entries:", len(manifest))
# checksums of the reference's own seeded (torch.manual_seed(42)) random initialisation, every tensor
init_sums = {k: float(v.double().sum()) for k, v in model.state_dict().items()}
with open(os.path.join(HERE, "seeded_init_sums.json"), "w") as f:
    json.dump(init_sums, f, indent=0, sort_keys=True)

# ---- BatchNorm calibration scalars (part of the weight definition, see tests/_weights.py) ----
CALIB_PATH = os.path.join(HERE, "bn_calibration.json")
if not os.path.exists(CALIB_PATH) or os.environ.get("EWVIT_RECALIBRATE"):
    fill_module_(model, seed=0, calib={})
    model.eval()
    calib = {}

    def make_hook(name):
        def hook(mod, inp):
            c = float(inp[0].float().std())
            c = float(f"{max(c, 1e-3):.3g}")
            calib[name] = c
            mod.running_var.mul_(c * c)
            mod.running_mean.mul_(c)
        return hook

    hooks = [m.register_forward_pre_hook(make_hook(n)) for n, m in model.named_modules()
             if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        cal_x = seeded_randn((2, 2, 3, 224, 224), 7)
        for mode in ("dynamic", "sfe_only", "sfe_mwt"):
            model(cal_x, batch_size=2, ablation=mode)
    for h in hooks:
        h.remove()
    with open(CALIB_PATH, "w") as f:
        json.dump(calib, f, indent=0, sort_keys=True)
    print("calibrated", len(calib), "BatchNorm layers")
    import _weights
    _weights._CALIB = None

fill_module_(model, seed=0)
model.eval()
gold = {"meta": {"torch": str(torch.__version__), "weights_seed": 0}}


def small(t):
    return t.detach().clone().contiguous()


with torch.no_grad():
    # ---- a-2: DWTForward(J=1,'haar','zero') as the reference calls it (mwt.py:76) ----
    dwt = DWTForward(J=1, wave="haar", mode="zero")
    for name, shape, seed in (("dwt_even", (2, 3, 32, 24), 11), ("dwt_odd", (2, 3, 9, 7), 12),
                              ("dwt_oddw", (1, 3, 8, 5), 13)):
        x = seeded_randn(shape, seed)
        ll, yh = dwt(x)
        gold[name] = {"x": small(x), "ll": small(ll), "yh": small(yh[0])}
    # three chained levels on a 224 frame: keep checksums + a crop (full tensors are MBs)
    x = seeded_randn((1, 3, 224, 224), 14)
    cur, rec = x, {}
    for lvl in range(3):
        ll, yh = dwt(cur)
        rec[f"ll{lvl+1}_sum"] = ll.double().sum()
        rec[f"ll{lvl+1}_abs"] = ll.double().abs().sum()
        rec[f"hf{lvl+1}_sum"] = yh[0].double().sum()
        rec[f"hf{lvl+1}_abs"] = yh[0].double().abs().sum()
        rec[f"ll{lvl+1}_crop"] = small(ll[:, :, :6, :6])
        rec[f"hf{lvl+1}_crop"] = small(yh[0][:, :, :, :6, :6])
        cur = ll
    gold["dwt3_224"] = rec

    # ---- a-3 / a-4: MWT ----
    frames = seeded_randn((2, 3, 224, 224), 21)
    gold["frames_seed"] = 21
    gold["frames_sum"] = frames.double().sum()
    mwt = model.dama.mwt
    ll, hfc = mwt.wavelet_transform(frames, (112, 112))
    gold["mwt_wt_l1"] = {"ll_crop": small(ll[:, :, :5, :5]), "hfc_crop": small(hfc[:, :8, :6, :6]),
                         "hfc_mean": small(hfc.mean(dim=(2, 3)))}
    ll2, hfc2 = mwt.wavelet_transform(ll, (112, 112))
    gold["mwt_wt_l2"] = {"hfc_crop": small(hfc2[:, :8, :6, :6]), "hfc_mean": small(hfc2.mean(dim=(2, 3)))}
    gold["mwt_out"] = small(mwt(frames))

    # ---- a-5 / a-6: SFE (V2-S backbone + 2-token ViT) ----
    feat = model.dama.sfe.efficient_net.features(frames)
    gold["backbone_feat_mean"] = small(feat.mean(dim=(2, 3)))
    gold["sfe_out"] = small(model.dama.sfe(frames))

    # ---- a-7 / a-8: DAMA._process_frame ----
    pf = model.dama._process_frame(frames)
    gold["process_frame"] = {k: small(v) for k, v in pf.items()}

    # ---- a-9 / a-10: full detector, ragged last chunk (K=5, batch_size=2) ----
    vids = seeded_randn((2, 5, 3, 224, 224), 31)
    out = model(vids, batch_size=2, ablation="dynamic")
    gold["detector_dynamic"] = {"seed": 31, "shape": [2, 5, 3, 224, 224], "batch_size": 2,
                                **{k: small(v) for k, v in out.items()}}
    # config 1 of BASELINE.json: x[1,8,3,224,224], batch_size=8
    vids1 = seeded_randn((1, 8, 3, 224, 224), 42)
    out1 = model(vids1, batch_size=8, ablation="dynamic")
    gold["detector_config1"] = {"seed": 42, "shape": [1, 8, 3, 224, 224], "batch_size": 8,
                                **{k: small(v) for k, v in out1.items()}}

    # ---- a-10: the two b0 ablation branches (model.py:100-161), ragged last chunk (K=5, batch_size=2) and K=4/bs=4 ----
    for tag, vv, bs in (("k5", vids, 2), ("k4", seeded_randn((2, 4, 3, 224, 224), 33), 4)):
        for mode in ("sfe_only", "sfe_mwt"):
            o = model(vv, batch_size=bs, ablation=mode)
            gold[f"detector_{mode}_{tag}"] = {"seed": 31 if tag == "k5" else 33, "shape": list(vv.shape), "batch_size": bs,
                                              **{k: (small(v) if torch.is_tensor(v) else v) for k, v in o.items()}}
            print(mode, tag, "logits:", o["logits"].flatten().tolist())
    # b0 backbone (third-party stand-in, see oracle/effnet_b0.py) and the b0-fed EfficientViT in both output modes
    b0 = model.sfe.efficient_net.extract_features(frames)
    gold["b0_feat_mean"] = small(b0.mean(dim=(2, 3)))
    gold["b0_feat_crop"] = small(b0[:, :16])
    gold["sfe_b0_out"] = small(model.sfe(frames))
    gold["sfe_cls_out"] = small(model.sfe_cls(frames))

    # ---- quirk (ii): more than 64 frames per chunk raises (sfe.py:158-159) ----
    try:
        model.dama.sfe(torch.zeros(65, 3, 224, 224))
        gold["n65_raises"] = False
    except RuntimeError as e:
        gold["n65_raises"] = True
        gold["n65_message"] = str(e)

    # ---- quirk (iv): fusion_gate 3x3 conv on a 1x1 map == its centre tap ----
    cat = seeded_randn((3, 256, 1, 1), 51)
    gold["fusion_gate"] = {"x": small(cat), "y": small(model.dama.fusion_gate(cat))}

torch.save(gold, os.path.join(HERE, "ewvit_golden.pt"))
print("wrote", os.path.join(HERE, "ewvit_golden.pt"), os.path.getsize(os.path.join(HERE, "ewvit_golden.pt")), "bytes")
print("logits (dynamic, K=5):", out["logits"].flatten().tolist())
print("logits (config1):", out1["logits"].flatten().tolist())
