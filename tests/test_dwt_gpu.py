"""GPU parity for the Haar kernels (row a-2): CUDA path through the C ABI vs the CPU oracle.
fp32, fixed evaluation order -> BIT-EXACT against oracle/haar.py (stricter than the 1e-5 relative
tolerance BASELINE.json's north_star asks for against the reference)."""
import numpy as np
import pytest
import torch

from oracle.haar import haar_dwt2, haar_dwt2_multilevel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from ewvit import ops
    return ops


def _randn(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("shape", [(1, 1, 2, 2), (2, 3, 8, 8), (1, 3, 9, 7), (3, 2, 5, 16), (2, 3, 224, 224),
                                   (1, 3, 1, 1), (1, 1, 1, 6), (1, 1, 33, 1)])
def test_single_level_matches_oracle_bitwise(ops, shape):
    x = _randn(shape, 1)
    ll, yh = ops.dwt_haar(x.cuda())
    ll_o, yh_o = haar_dwt2(x)
    assert torch.equal(ll.cpu(), ll_o) and torch.equal(yh.cpu(), yh_o)


def test_single_level_empty_input(ops):
    ll, yh = ops.dwt_haar(torch.zeros(0, 3, 8, 8, device="cuda"))
    assert ll.shape == (0, 3, 4, 4) and yh.shape == (0, 3, 3, 4, 4)


def test_single_level_golden_reference_call(ops, golden):
    for name in ("dwt_even", "dwt_odd", "dwt_oddw"):
        g = golden[name]
        ll, yh = ops.dwt_haar(g["x"].cuda())
        assert (ll.cpu() - g["ll"]).abs().max() <= 1e-5 * g["ll"].abs().max()
        assert (yh.cpu() - g["yh"]).abs().max() <= 1e-5 * g["yh"].abs().max()


@pytest.mark.parametrize("shape", [(1, 1, 8, 8), (2, 3, 224, 224), (5, 3, 64, 48), (1, 3, 8, 2048), (7, 1, 40, 8),
                                   (64, 3, 224, 224)])
def test_fused_three_levels_match_oracle_bitwise(ops, shape):
    x = _randn(shape, 2)
    out = ops.dwt3_haar(x.cuda())
    for lvl, (ll, yh) in enumerate(haar_dwt2_multilevel(x, 3), start=1):
        assert torch.equal(out[f"ll{lvl}"].cpu(), ll), f"ll{lvl}"
        assert torch.equal(out[f"hf{lvl}"].cpu(), yh), f"hf{lvl}"


def test_fused_skips_null_outputs(ops):
    x = _randn((2, 3, 32, 32), 3)
    out = ops.dwt3_haar(x.cuda(), want=("hf1", "hf2", "hf3"))
    assert sorted(out) == ["hf1", "hf2", "hf3"]
    ref = haar_dwt2_multilevel(x, 3)
    for lvl in (1, 2, 3):
        assert torch.equal(out[f"hf{lvl}"].cpu(), ref[lvl - 1][1])


def test_fused_rejects_ragged(ops):
    from ewvit import EwvitError
    with pytest.raises(EwvitError):
        ops.dwt3_haar(torch.zeros(1, 3, 12, 16, device="cuda"))


def test_fused_equals_three_single_levels(ops):
    x = _randn((3, 3, 224, 224), 4).cuda()
    out = ops.dwt3_haar(x)
    cur = x
    for lvl in (1, 2, 3):
        ll, yh = ops.dwt_haar(cur)
        assert torch.equal(ll, out[f"ll{lvl}"]) and torch.equal(yh, out[f"hf{lvl}"])
        cur = ll


def test_config2_full_size_properties(ops):
    """BASELINE.json configs[1]: 256x3x224x224.  Bit-compare against the oracle plus the
    size-independent Haar identities (energy conservation per level, LL3 of a constant = 8c)."""
    x = _randn((256, 3, 224, 224), 42)
    out = ops.dwt3_haar(x.cuda())
    ref = haar_dwt2_multilevel(x, 3)
    for lvl in (1, 2, 3):
        assert torch.equal(out[f"ll{lvl}"].cpu(), ref[lvl - 1][0])
        assert torch.equal(out[f"hf{lvl}"].cpu(), ref[lvl - 1][1])
    e = x.double().pow(2).sum()
    e3 = out["ll3"].double().pow(2).sum() + sum(out[f"hf{l}"].double().pow(2).sum() for l in (1, 2, 3))
    assert abs(e3.item() - e.item()) / e.item() < 1e-6
    c = ops.dwt3_haar(torch.full((4, 3, 224, 224), 0.25, device="cuda"))
    assert torch.allclose(c["ll3"], torch.full_like(c["ll3"], 2.0), rtol=1e-6)
    assert all(float(c[f"hf{l}"].abs().max()) == 0.0 for l in (1, 2, 3))


def test_fused_three_levels_from_uint8_frames_bitwise(ops):
    """uint8 input + on-load ToTensor/Normalize (row f-3) == the fp32 kernel on torchvision-normalised frames, bit for bit"""
    g = torch.Generator().manual_seed(7)
    u = torch.randint(0, 256, (5, 3, 64, 48), generator=g, dtype=torch.uint8)
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])
    xf = u.float().div(255).sub(mean.view(1, 3, 1, 1)).div(std.view(1, 3, 1, 1))        # to_tensor + normalize
    ref = ops.dwt3_haar(xf.cuda())
    got = ops.dwt3_haar(u.cuda(), norm=(mean.cuda(), std.cuda()))
    for k in ref:
        assert torch.equal(got[k], ref[k]), k
    ll, hf = haar_dwt2(xf)
    assert torch.equal(got["ll1"].cpu(), ll) and torch.equal(got["hf1"].cpu(), hf)
