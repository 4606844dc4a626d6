"""Stand-in for the third-party ``pytorch_wavelets`` package.  TEST INFRASTRUCTURE ONLY.

Lets the UNMODIFIED reference ``network/mwt.py`` (which does
``from pytorch_wavelets import DWTForward`` at line 5) import in the build container,
where the real package is absent and cannot be installed (no network).  Only
``tests/golden/make_golden.py`` puts this directory on ``sys.path``.

It restates the published algorithm of ``pytorch_wavelets.DWTForward`` for the one
configuration the reference uses (``J=1, wave='haar', mode='zero'``, mwt.py:20) as the
package itself does it -- two grouped stride-2 ``F.conv2d`` passes -- so that the
slicing oracle in ``oracle/haar.py`` has an independent formulation to be compared with.
PARITY UNPINNED: neither formulation is checked against the real package.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

_HAAR_DEC_LO = (0.7071067811865476, 0.7071067811865476)    # pywt.Wavelet('haar').dec_lo
_HAAR_DEC_HI = (-0.7071067811865476, 0.7071067811865476)   # pywt.Wavelet('haar').dec_hi


def _analysis_1d(x, lo, hi, dim):
    """Stride-2 grouped cross-correlation with [lo, hi] per channel along ``dim`` (2=H, 3=W),
    zero mode: an odd length gets one zero appended at the end, nothing else."""
    c = x.shape[1]
    n = x.shape[dim]
    taps = lo.numel()
    out = (n + taps - 1) // 2
    p = 2 * (out - 1) - n + taps
    if p % 2 == 1:
        x = F.pad(x, (0, 0, 0, 1) if dim == 2 else (0, 1, 0, 0))
    bank = torch.cat([lo, hi] * c, dim=0)
    if dim == 2:
        return F.conv2d(x, bank, stride=(2, 1), padding=(p // 2, 0), groups=c)
    return F.conv2d(x, bank, stride=(1, 2), padding=(0, p // 2), groups=c)


class DWTForward(nn.Module):
    def __init__(self, J=1, wave="haar", mode="zero"):
        super().__init__()
        if wave != "haar" or mode != "zero":
            raise NotImplementedError("shim covers the reference's configuration only (mwt.py:20)")
        lo = torch.tensor(_HAAR_DEC_LO[::-1], dtype=torch.float32)   # filters are stored reversed
        hi = torch.tensor(_HAAR_DEC_HI[::-1], dtype=torch.float32)
        self.register_buffer("h0_col", lo.reshape(1, 1, -1, 1).clone())
        self.register_buffer("h1_col", hi.reshape(1, 1, -1, 1).clone())
        self.register_buffer("h0_row", lo.reshape(1, 1, 1, -1).clone())
        self.register_buffer("h1_row", hi.reshape(1, 1, 1, -1).clone())
        self.J = J
        self.mode = mode

    def forward(self, x):
        highs = []
        low = x
        for _ in range(self.J):
            rows = _analysis_1d(low, self.h0_row, self.h1_row, 3)
            both = _analysis_1d(rows, self.h0_col, self.h1_col, 2)
            n, _, h, w = both.shape
            both = both.reshape(n, -1, 4, h, w)
            low = both[:, :, 0].contiguous()
            highs.append(both[:, :, 1:].contiguous())
        return low, highs
