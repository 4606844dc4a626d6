"""SE gate kernel at the V2-S shapes, 512 frames (probe tool)."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
for (c, sq) in ((256, 16), (512, 32), (768, 32), (960, 40), (1536, 64)):
    n = 512
    pooled = torch.randn(n, c, device="cuda")
    w1, b1 = torch.randn(sq, c, device="cuda") * c ** -0.5, torch.randn(sq, device="cuda")
    w2t, b2 = torch.randn(sq, c, device="cuda") * sq ** -0.5, torch.randn(c, device="cuda")
    out = torch.empty(n, c, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.se_gate(pooled, w1, b1, w2t, b2, out=out, bf16=True)
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(20):
                fn()
    torch.cuda.current_stream().wait_stream(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"c={c:5d} sq={sq:3d}: {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us per launch")
