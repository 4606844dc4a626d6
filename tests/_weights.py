"""Deterministic, key-addressed weight filler shared by the golden-vector generator and the tests.

Weights are a pure function of (state_dict key, shape, seed): no dependence on module
construction order or on the global RNG, so the build container (reference modules) and the
GPU box (this repo's modules / the functional oracle) get bit-identical tensors.
"""
import json
import os
import zlib

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


_CALIB = None


def bn_calibration():
    """Per-BatchNorm scalar c (committed, ``golden/bn_calibration.json``): the filler's running
    statistics are multiplied by c (mean) / c*c (var) so that activations stay O(1) through the
    ~110 BatchNorm layers of the backbones instead of growing by orders of magnitude.  The
    scalars were measured once by ``make_golden.py``; they are part of the weight DEFINITION."""
    global _CALIB
    if _CALIB is None:
        path = os.path.join(GOLDEN_DIR, "bn_calibration.json")
        _CALIB = json.load(open(path)) if os.path.exists(path) else {}
    return _CALIB


def fill_tensor(key: str, shape, dtype, seed: int = 0, calib=None):
    """None means: leave the entry alone (the fixed Haar filter buffers)."""
    shape = tuple(shape)
    calib = bn_calibration() if calib is None else calib
    c = float(calib.get(key.rsplit(".", 1)[0], 1.0))
    if ".dwt." in key or key.startswith("dwt."):
        return None
    g = _gen(key, seed)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.int64)
    if key.endswith("running_var"):
        return (torch.rand(shape, generator=g) + 0.5) * (c * c)
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * (0.1 * c)
    if key.endswith("cls_token") or key.endswith("pos_embedding"):
        return torch.randn(shape, generator=g) * 0.5
    if len(shape) <= 1:
        if key.endswith("weight"):
            return torch.rand(shape, generator=g) * 0.4 + 0.8
        return torch.randn(shape, generator=g) * 0.05
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 1.4142135 if len(shape) == 4 else 1.0
    return torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)


def load_manifest(name="state_dict_manifest.json"):
    with open(os.path.join(GOLDEN_DIR, name)) as f:
        return json.load(f)


def state_dict_from_manifest(manifest, seed: int = 0, prefixes=None):
    """Build a full fp32 state_dict {key: tensor} for the keys under ``prefixes``."""
    sd = {}
    for key, meta in manifest.items():
        if prefixes is not None and not any(key.startswith(p) for p in prefixes):
            continue
        t = fill_tensor(key, meta["shape"], meta["dtype"], seed)
        if t is None:
            t = haar_buffer(key)
        sd[key] = t
    return sd


def haar_buffer(key: str):
    s = 0.7071067811865476
    lo, hi = torch.tensor([s, s], dtype=torch.float32), torch.tensor([s, -s], dtype=torch.float32)
    v = lo if ".h0_" in key else hi
    return v.reshape(1, 1, 2, 1).clone() if key.endswith("_col") else v.reshape(1, 1, 1, 2).clone()


def fill_module_(module: torch.nn.Module, seed: int = 0, prefix: str = "", calib=None):
    """In-place: overwrite every parameter and buffer of ``module`` with the filler."""
    with torch.no_grad():
        for key, t in module.state_dict().items():
            v = fill_tensor(prefix + key, t.shape, str(t.dtype), seed, calib)
            if v is not None:
                t.copy_(v.to(t.dtype))
    return module


def seeded_randn(shape, seed: int):
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randn(shape, generator=g)
