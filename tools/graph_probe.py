"""Probe: does replaying the whole 512-frame forward as ONE CUDA graph beat eager launches?  (inter-kernel gaps / host overhead)"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from network.model import DeepfakeDetector  # noqa: E402

torch.manual_seed(42)
m = DeepfakeDetector(3, 128, batch_size=8).cuda().eval()
x = torch.randn(8, 64, 3, 224, 224, device="cuda")


def timeit(fn, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    for _ in range(5):
        ref = m(x, 8, "dynamic")["logits"].clone()
    print(f"eager : {timeit(lambda: m(x, 8, 'dynamic')):.3f} ms/step")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        m(x, 8, "dynamic")
        with torch.cuda.graph(g, stream=side):
            out = m(x, 8, "dynamic")
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    print("graph == eager:", bool(torch.equal(out["logits"], ref)))
    print(f"graph : {timeit(g.replay):.3f} ms/step")
    print(f"eager : {timeit(lambda: m(x, 8, 'dynamic')):.3f} ms/step")
