"""Per-layer HBM roofline of the native EfficientNetV2-S at n frames vs measured (reads a backbone_bench.py listing)."""
import ast
import sys

n = 512
BW = 6.5456e12
lines = open(sys.argv[1]).read().splitlines()[1:]
# spatial size per layer index: replay the network
hw = 112
rows = []
cin_prev = 24
tot_m = tot_r = 0.0
agg = {}
for ln in lines:
    name, ms, shp = ast.literal_eval(ln)
    idx, kind = int(name.split(".")[1]), name.split(".")[2]
    rows.append((idx, kind, ms, shp[0]))
# geometry by torchvision V2-S stage table
stages = [("f", 1, 24, 2, 1), ("f", 4, 48, 4, 2), ("f", 4, 64, 4, 2), ("m", 4, 128, 6, 2), ("m", 6, 160, 9, 1), ("m", 6, 256, 15, 2)]
geo = {}
i = 1
cin = 24
hw = 112
for typ, e, cout, reps, s in stages:
    for r in range(reps):
        st = s if r == 0 else 1
        ci = cin if r == 0 else cout
        ho = hw // st
        if typ == "f":
            if e == 1:
                geo[i] = (hw, ho, ci, cout, 9); i += 1
            else:
                geo[i] = (hw, ho, ci, ci * e, 9); i += 1
                geo[i] = (ho, ho, ci * e, cout, 1); i += 1
        else:
            geo[i] = (hw, hw, ci, ci * e, 1); i += 1          # expand
            geo[i] = (hw, ho, ci * e, ci * e, "dw"); i += 1   # dw
            geo[i] = (ho, ho, ci * e, ci * e, "se"); i += 1   # se
            geo[i] = (ho, ho, ci * e, cout, 1); i += 1        # project
        hw = ho
    cin = cout
geo[i] = (7, 7, 256, 1280, 1)
for idx, kind, ms, shp in rows:
    if idx == 0:
        by = n * (3 * 224 * 224 * 4 + 112 * 112 * 24 * 2); fl = 0
    else:
        hi, ho, ci, co, k = geo[idx]
        if k == "se":
            by = 0 if kind == "se" and True else 0
            fl = 0
        elif k == "dw":
            by = n * (hi * hi * ci + ho * ho * co) * 2; fl = n * ho * ho * co * 9 * 2
        else:
            by = n * (hi * hi * ci + ho * ho * co) * 2; fl = n * ho * ho * co * ci * k * 2
    t_roof = max(by / BW, fl / 1.386e15) * 1e3
    agg.setdefault(kind, [0.0, 0.0])
    agg[kind][0] += ms; agg[kind][1] += t_roof
    tot_m += ms; tot_r += t_roof
    if len(sys.argv) > 2:
        print(f"{idx:3d} {kind:7s} {str(shp):14s} meas {ms*1e3:7.1f} us  roof {t_roof*1e3:7.1f} us  x{ms/max(t_roof,1e-9):5.2f}")
for k, (m, r) in agg.items():
    print(f"{k:7s} measured {m:6.3f} ms   roofline {r:6.3f} ms   x{m/max(r,1e-9):.2f}")
print(f"total   measured {tot_m:6.3f} ms   roofline {tot_r:6.3f} ms")
