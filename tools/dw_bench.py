"""Depthwise kernels at the V2-S shapes (512 frames): old TMA-staged kernel vs ewvit_dwconv_nhwc_bf16 (probe tool)."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

n = 512
for (c, s, hw) in ((256, 2, 28), (512, 1, 14), (768, 1, 14), (960, 1, 14), (960, 2, 14), (1536, 1, 7)):
    # several distinct input buffers so consecutive launches do not hit in L2 (the real pipeline streams from the previous layer)
    xs = [torch.randn(n, hw, hw, c, device="cuda").bfloat16() for _ in range(4)]
    w = torch.randn(9, c, device="cuda") / 3
    b = torch.randn(c, device="cuda") * 0.1
    ho = (hw - 1) // s + 1
    y = torch.empty(n, ho, ho, c, device="cuda", dtype=torch.bfloat16)
    pooled = torch.empty(n, c, device="cuda")
    it = [0]
    def old():
        it[0] += 1
        ops.dwconv3x3(xs[it[0] % 4], w, b, s, out=y, pooled=pooled)
    def new():
        it[0] += 1
        ops.dwconv(xs[it[0] % 4], w, b, 3, s, out=y, pooled=True)
    byt = (n * hw * hw * c + n * ho * ho * c) * 2
    to = timeit(old) if hasattr(ops, "dwconv3x3") and os.environ.get("EWVIT_DW_OLD", "1") == "1" else float("nan")
    tn = timeit(new)
    print(f"c={c:5d} s={s} hw={hw:3d}  old {to:7.1f} us ({byt / to / 1e6 if to == to else 0:5.0f} GB/s)   new {tn:7.1f} us ({byt / tn / 1e6:5.0f} GB/s)  "
          f"[variant {os.environ.get('EWVIT_DW_VARIANT', '0')} ws7={os.environ.get('EWVIT_DW_WS7', '0')}]")
