"""Quick CUDA-event timings of the tcgen05 conv/GEMM kernel at the model's real shapes (probe tool)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    e[0].record()
    for i in range(iters):
        fn()
        e[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(iters))
    return ts[len(ts) // 2]


res = {}
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = "cuda"
for name, cin in (("hf_fusion_64to128", 64), ("multiscale_384to128", 384)):
    x = torch.randn(nf, 114, 114, cin, device=dev).bfloat16()
    w = (torch.randn(128, 3, 3, cin, device=dev) * (9 * cin) ** -0.5).bfloat16()
    y = torch.empty(nf, 114, 114, 128, device=dev, dtype=torch.bfloat16)
    sc, sh = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    ms = timeit(lambda: ops.conv3x3_bf16(x, w, nf, 112, 112, 1, True, sc, sh, True, y, 0, True))
    flops = 2.0 * nf * 112 * 112 * 128 * 9 * cin
    res[name] = {"ms": ms, "TFLOPs_useful": flops / ms / 1e9, "frames": nf}
x = torch.randn(nf, 114, 114, 128, device=dev).bfloat16()
w = (torch.randn(128, 3, 3, 128, device=dev) * (9 * 128) ** -0.5).bfloat16()
y = torch.empty(nf, 56, 56, 128, device=dev, dtype=torch.bfloat16)
sc, sh = torch.ones(128, device=dev), torch.zeros(128, device=dev)
ms = timeit(lambda: ops.conv3x3_bf16(x, w, nf, 112, 112, 2, True, sc, sh, True, y, 0, False))
res["freq_conv_s2"] = {"ms": ms, "TFLOPs_useful": 2.0 * nf * 56 * 56 * 128 * 9 * 128 / ms / 1e9}
for name, (m, n, k, splits) in {"patch_embed": (512, 512, 62720, 9), "vit_qkv": (1024, 1536, 512, 1),
                                "vit_ff1": (1024, 2048, 512, 1), "vit_ff2": (1024, 512, 2048, 1),
                                "big_8192": (8192, 8192, 8192, 1)}.items():
    a = torch.randn(m, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
    out = torch.empty(m, n, device=dev)
    ws = torch.empty(splits, m, n, device=dev) if splits > 1 else None
    ms = timeit(lambda: ops.linear_bf16(a, w, out=out, splits=splits, workspace=ws))
    ref_ms = timeit(lambda: torch.matmul(a, w.t()))
    res[name] = {"ms": ms, "TFLOPs": 2.0 * m * n * k / ms / 1e9, "torch_matmul_ms": ref_ms}
print(json.dumps(res, indent=1))
