"""Oracle Haar DWT: analytic known answers + the reference-call golden vectors (a-2)."""
import numpy as np
import torch

from oracle.haar import HAAR_S, haar_dwt2, haar_dwt2_np, haar_dwt2_multilevel


def test_constant_image():
    x = np.full((1, 1, 8, 8), 3.0, np.float32)
    ll, yh = haar_dwt2_np(x)
    assert np.allclose(ll, 6.0, rtol=1e-6)          # LL = 2c
    assert np.all(yh == 0)


def test_three_level_constant():
    x = torch.full((1, 3, 224, 224), 1.5)
    levels = haar_dwt2_multilevel(x, 3)
    assert torch.allclose(levels[2][0], torch.full((1, 3, 28, 28), 12.0), rtol=1e-6)   # 8c
    assert all((yh == 0).all() for _, yh in levels)


def test_impulses_and_signs():
    # a=x[0,0], b=x[0,1], c=x[1,0], d=x[1,1];  LH=(a+b-c-d)/2  HL=(a-b+c-d)/2  HH=(a-b-c+d)/2
    signs = {(0, 0): (1, 1, 1), (0, 1): (1, -1, -1), (1, 0): (-1, 1, -1), (1, 1): (-1, -1, 1)}
    for (i, j), sg in signs.items():
        x = np.zeros((1, 1, 2, 2), np.float32)
        x[0, 0, i, j] = 1.0
        ll, yh = haar_dwt2_np(x)
        assert np.isclose(ll[0, 0, 0, 0], 0.5, rtol=1e-6)
        for k in range(3):
            assert np.isclose(yh[0, 0, k, 0, 0], 0.5 * sg[k], rtol=1e-6)


def test_checkerboard_only_hh():
    x = np.indices((8, 8)).sum(0) % 2 * 2.0 - 1.0
    ll, yh = haar_dwt2_np(x[None, None].astype(np.float32))
    assert np.all(ll == 0) and np.all(yh[:, :, 0] == 0) and np.all(yh[:, :, 1] == 0)
    assert np.allclose(np.abs(yh[:, :, 2]), 2.0, rtol=1e-6)


def test_parseval_and_reconstruction():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 3, 16, 12)).astype(np.float32)
    ll, yh = haar_dwt2_np(x)
    e_in = (x.astype(np.float64) ** 2).sum()
    e_out = (ll.astype(np.float64) ** 2).sum() + (yh.astype(np.float64) ** 2).sum()
    assert abs(e_in - e_out) / e_in < 1e-6
    lh, hl, hh = yh[:, :, 0], yh[:, :, 1], yh[:, :, 2]
    rec = np.empty_like(x)
    rec[..., 0::2, 0::2] = (ll + lh + hl + hh) / 2
    rec[..., 0::2, 1::2] = (ll + lh - hl - hh) / 2
    rec[..., 1::2, 0::2] = (ll - lh + hl - hh) / 2
    rec[..., 1::2, 1::2] = (ll - lh - hl + hh) / 2
    assert np.allclose(rec, x, atol=1e-6)


def test_fp32_scale_is_not_exact_half():
    assert float(HAAR_S * HAAR_S) != 0.5


def test_numpy_and_torch_twins_bit_identical():
    x = torch.randn(2, 3, 10, 14, generator=torch.Generator().manual_seed(3))
    ll_t, yh_t = haar_dwt2(x)
    ll_n, yh_n = haar_dwt2_np(x.numpy())
    assert np.array_equal(ll_t.numpy(), ll_n) and np.array_equal(yh_t.numpy(), yh_n)


def test_zero_mode_odd_sizes_pad_bottom_right():
    x = np.arange(1, 16, dtype=np.float32).reshape(1, 1, 3, 5)
    ll, yh = haar_dwt2_np(x)
    assert ll.shape == (1, 1, 2, 3) and yh.shape == (1, 1, 3, 2, 3)
    xp = np.zeros((1, 1, 4, 6), np.float32)
    xp[..., :3, :5] = x
    ll2, yh2 = haar_dwt2_np(xp)
    assert np.array_equal(ll, ll2) and np.array_equal(yh, yh2)


def test_against_reference_call_golden(golden):
    """DWTForward(J=1,'haar','zero') outputs produced through the reference's call site."""
    for name in ("dwt_even", "dwt_odd", "dwt_oddw"):
        g = golden[name]
        ll, yh = haar_dwt2(g["x"])
        assert ll.shape == g["ll"].shape and yh.shape == g["yh"].shape
        scale = g["yh"].abs().max()
        assert (ll - g["ll"]).abs().max() <= 1e-6 * g["ll"].abs().max()
        assert (yh - g["yh"]).abs().max() <= 1e-6 * scale


def test_three_levels_224_golden(golden):
    from _weights import seeded_randn
    g = golden["dwt3_224"]
    levels = haar_dwt2_multilevel(seeded_randn((1, 3, 224, 224), 14), 3)
    for i, (ll, yh) in enumerate(levels, start=1):
        assert torch.allclose(ll[:, :, :6, :6], g[f"ll{i}_crop"], rtol=0, atol=1e-6 * float(ll.abs().max()))
        assert torch.allclose(yh[:, :, :, :6, :6], g[f"hf{i}_crop"], rtol=0, atol=1e-6 * float(yh.abs().max()))
        assert abs(ll.double().abs().sum() - g[f"ll{i}_abs"]) <= 1e-6 * g[f"ll{i}_abs"]
        assert abs(yh.double().abs().sum() - g[f"hf{i}_abs"]) <= 1e-6 * g[f"hf{i}_abs"]
