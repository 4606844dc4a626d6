"""Three-level MWT head at 112x112: upsample3 and head_conv3 timed separately, role timeline of CTA 0 (probe tool)."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "efficient-wavelet-vit_b200"))
from ewvit import ops  # noqa: E402
from ewvit._lib import load  # noqa: E402
lib = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hfs = [torch.randn(n, 9, 112 >> l, 112 >> l, device="cuda") for l in range(3)]
up = torch.zeros(n, 114, 114, 32, device="cuda", dtype=torch.bfloat16)
w = (torch.randn(128, 288, device="cuda") * 0.1).bfloat16()
y = torch.empty(n, 114, 114, 192, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(192, device="cuda"), torch.zeros(192, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(lambda: ops.mwt_upsample3(*hfs, up, 112, 112))
print(f"upsample3: {ms * 1e3:.1f} us ({(sum(t.numel() for t in hfs) * 4 + up.numel() * 2) / ms / 1e6:.0f} GB/s)")
fn = lambda: ops.mwt_head_conv3(up, w, sc, sh, y, 112, 112)
for flags in (0, 512):
    lib.ewvit_debug_set_flags(flags)
    ms = timeit(fn)
    print(f"=== head_conv3 flags {flags}: {ms * 1e3:.1f} us ({(up.numel() + y.numel()) * 2 / ms / 1e6:.0f} GB/s)")
    buf = torch.zeros(8 * 64 * 4, dtype=torch.int64, device="cuda")
    lib.ewvit_debug_set_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.ewvit_debug_set_trace(None)
    t = buf.cpu().view(8, 64, 4)
    t0 = int(t[t > 0].min())
    for tile in range(20, 26):
        row = []
        for role, nm in enumerate(("tma", "mma", "epi0", "epi1", "epi2", "epi3")):
            v = t[role, tile]
            if int(v.max()) == 0:
                continue
            row.append(f"{nm}:" + ",".join(str(int(x) - t0) if int(x) else "-" for x in v))
        print(f"tile {tile:2d}  " + "  ".join(row))
lib.ewvit_debug_set_flags(0)
# the fusion conv reading one level in place vs a dense 64-channel tensor
x64 = torch.randn(n, 114, 114, 64, device="cuda").bfloat16()
wf = (torch.randn(128, 3, 3, 64, device="cuda") * 0.05).bfloat16()
cat = torch.empty(n, 114, 114, 384, device="cuda", dtype=torch.bfloat16)
s128, z128 = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
ms = timeit(lambda: ops.conv3x3_bf16(x64, wf, n, 112, 112, 1, True, s128, z128, True, cat, 0, True))
print(f"fusion conv, dense 64-channel input: {ms * 1e3:.1f} us")
ms = timeit(lambda: ops.conv3x3_bf16(y, wf, n, 112, 112, 1, True, s128, z128, True, cat, 128, True, x_coff=64))
print(f"fusion conv, level slice of the 192-channel head: {ms * 1e3:.1f} us")
