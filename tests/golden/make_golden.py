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

# ---- f-2: ONE TRAINING MICRO-STEP of the unmodified reference (train.py:93-111 semantics): train mode, BatchNorm batch
#      statistics per chunk of B*batch_size frames (dama.py:179-186), every Dropout / StochasticDepth probability set to 0 so the
#      step is deterministic, loss = BinaryFocalLoss (config/focal_loss.py) + orthogonal loss at lambda = 1 (train.py:55-91)
from config.focal_loss import BinaryFocalLoss  # noqa: E402  (the reference's own module)

fill_module_(model, seed=0)
model.train()
for m in model.modules():
    if isinstance(m, torch.nn.Dropout) or type(m).__name__ == "StochasticDepth":
        m.p = 0.0
tx = seeded_randn((2, 4, 3, 224, 224), 71)
ty = torch.tensor([0.0, 1.0])
model.zero_grad(set_to_none=True)
tout = model(tx, batch_size=2, ablation="dynamic")


def _orth(space, freq):                                   # train.py:55-67
    d = space.shape[1]
    sp, fq = torch.nn.functional.normalize(space, p=2, dim=1), torch.nn.functional.normalize(freq, p=2, dim=1)
    off = torch.mm(sp.T, fq) * (1 - torch.eye(d))
    return torch.norm(off, p="fro") ** 2 / (d * (d - 1))


cls_loss = BinaryFocalLoss()(tout["logits"], ty.view(-1, 1))
orth_loss = _orth(tout["space"], tout["freq"])
(cls_loss + orth_loss).backward()
named = dict(model.named_parameters())
GRAD_KEYS = ["classifier.3.weight", "classifier.0.bias", "dama.gate_net.5.weight", "dama.cross_att.layers.1.3.to_q.weight",
             "dama.fusion_gate.0.bias", "dama.sfe.feat_map.0.weight", "dama.sfe.patch_to_embedding.bias",
             "dama.sfe.transformer.layers.0.0.fn.to_qkv.weight", "dama.mwt.freq_pool.1.weight", "dama.mwt.multiscale_fusion.0.bias",
             "dama.mwt.hf_conv.seperate.1.0.weight", "dama.sfe.efficient_net.features.7.0.weight",
             "dama.sfe.efficient_net.features.2.0.block.0.0.weight"]
BUF_KEYS = ["dama.mwt.freq_conv.1.running_mean", "dama.mwt.freq_conv.1.running_var", "dama.mwt.freq_conv.1.num_batches_tracked",
            "dama.fusion_gate.1.running_var", "dama.sfe.efficient_net.features.7.1.running_mean"]
sdt = model.state_dict()
gold["train_step"] = {
    "seed": 71, "shape": [2, 4, 3, 224, 224], "batch_size": 2, "labels": ty.clone(),
    "logits": small(tout["logits"]), "fused": small(tout["fused"]), "space": small(tout["space"]), "freq": small(tout["freq"]),
    "cls_loss": float(cls_loss), "orth_loss": float(orth_loss),
    "grads": {k: small(named[k].grad if named[k].grad.numel() <= 4096 else named[k].grad.flatten()[:4096]) for k in GRAD_KEYS},
    "grad_norms": {k: float(named[k].grad.norm()) for k in GRAD_KEYS},
    "no_grad": [k for k in ("mwt.freq_conv.0.weight", "sfe.feat_map.0.weight", "sfe_cls.mlp_head.0.weight", "fusion_gate.0.weight")
                if named[k].grad is None],
    "buffers": {k: small(sdt[k]) for k in BUF_KEYS},
}
print("train step: cls", float(cls_loss), "orth", float(orth_loss), "logits", tout["logits"].flatten().tolist())
model.eval()

torch.save(gold, os.path.join(HERE, "ewvit_golden.pt"))
print("wrote", os.path.join(HERE, "ewvit_golden.pt"), os.path.getsize(os.path.join(HERE, "ewvit_golden.pt")), "bytes")
print("logits (dynamic, K=5):", out["logits"].flatten().tolist())
print("logits (config1):", out1["logits"].flatten().tolist())
