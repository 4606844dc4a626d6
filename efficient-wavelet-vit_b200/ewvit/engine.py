"""Host-side runners of the native EWViT forward path (eval mode, bf16 tensor-core math, fp32 glue).

Each runner owns (a) weights re-laid-out for the kernels (folded BatchNorm, tap-major bf16 conv weights,
transposed DAMA matrices) and (b) per-batch-size workspaces, and strings the C-ABI kernels of
``libewvit.so`` together.  PyTorch is used for memory, streams and the third-party EfficientNet backbone
(cuDNN, bf16 channels-last) only.  Nothing here runs on the CPU and nothing falls back.

Reference call stacks being replaced: ``MWT.forward`` (network/mwt.py:92-119), ``EfficientViT.forward``
(network/sfe.py:145-173), ``DAMA._process_frame``/``DAMA.forward`` (network/dama.py:130-206) and the
``dynamic`` branch of ``DeepfakeDetector.forward`` (network/model.py:83-99).
"""
import copy
import math

import torch
from torch import nn

from . import ops
from ._lib import EwvitError

BN_EPS = 1e-5
LN_EPS = 1e-5
MACRO_BATCH = 512          # frames per pass: bounds the MWT workspace (~22 MB / frame)


class StageTimer:
    """Optional CUDA-event bracketing of the pipeline stages (bench.py turns it on; off by default).
    Events are recorded on the current stream, so a bracket measures exactly the kernels launched inside it."""

    def __init__(self):
        self.records = {}

    class _Span:
        def __init__(self, timer, name):
            self.timer, self.name = timer, name

        def __enter__(self):
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

        def __exit__(self, *exc):
            self.e1.record()
            self.timer.records.setdefault(self.name, []).append((self.e0, self.e1))

    def span(self, name):
        return StageTimer._Span(self, name)

    def summary_ms(self):
        """name -> (mean ms per bracket, brackets).  Call after a device synchronize."""
        return {k: (sum(a.elapsed_time(b) for a, b in v) / len(v), len(v)) for k, v in self.records.items()}

    def reset(self):
        self.records = {}


class _NullSpan:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullSpan()
TIMER = None               # set to a StageTimer to collect per-stage device times


def stage(name):
    return TIMER.span(name) if TIMER is not None else _NULL


def _fold_bn(sd, conv, bn, eps=BN_EPS):
    """conv bias + eval BatchNorm -> per-channel (scale, shift) applied to the bias-free conv output."""
    scale = sd[bn + "weight"].float() / torch.sqrt(sd[bn + "running_var"].float() + eps)
    shift = sd[bn + "bias"].float() + (sd[conv + "bias"].float() - sd[bn + "running_mean"].float()) * scale
    return scale.contiguous(), shift.contiguous()


def _conv_w_tapmajor(w, cin_pad=None):
    """[cout, cin, 3, 3] -> bf16 [cout, 3, 3, cin_pad] (k = (ky*3+kx)*cin_pad + c)."""
    cout, cin = w.shape[:2]
    cin_pad = cin_pad or cin
    out = torch.zeros((cout, 3, 3, cin_pad), dtype=torch.bfloat16, device=w.device)
    out[..., :cin] = w.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.contiguous()


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _host_sd(sd):
    """(CPU copies of the state-dict tensors, device they live on).  Every runner folds / re-lays-out its weights on the
    HOST (a few ms of fp32 CPU math on <= 61 M parameters) and uploads the results once: a rebuild after load_state_dict
    or an optimizer step costs two memcpys per tensor instead of ~10 tiny elementwise launches per layer."""
    dev = next(iter(sd.values())).device
    return {k: v.detach().to("cpu") for k, v in sd.items()}, dev


def _to_device(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: _to_device(v, dev) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_device(v, dev) for v in obj)
    return obj


# ------------------------------------------------------------------------------------------ MWT
def pack_head3_weights(wbd):
    """Block-diagonal head weights [64, 3, 3, 16] fp32 (output g*18+oc, tap, subband 3g+ic) -> bf16 [128, 288] for
    ``ewvit_mwt_head_conv3_fwd``: a pixel of its input holds channel 9*l + c for level l, subband c, i.e. two K steps of 16
    channels; per tap and K step one [128, 16] tile whose row halves are the shared weights of the two levels living in that
    step (step 0: levels 0 | 1, step 1: levels 1 | 2), placed at the positions their subbands occupy, zero elsewhere."""
    out = torch.zeros((2, 64, 9, 2, 16), dtype=torch.float32)      # [row half, out, tap, step, kk]
    w9 = wbd.reshape(64, 9, 16)[:, :, :9]                          # [out, tap, subband]
    for step in range(2):
        for half in range(2):
            lvl = step + half
            for kk in range(16):
                c = step * 16 + kk - 9 * lvl
                if 0 <= c < 9:
                    out[half, :, :, step, kk] = w9[:, :, c]
    return out.reshape(128, 288).to(torch.bfloat16).contiguous()


class MwtRunner:
    """Native ``MWT.forward`` (levels = 3, in_channels = 3).  ``sd`` holds the module's own keys
    (``hf_conv.seperate.0.0.weight`` ...) as CUDA tensors."""

    def __init__(self, sd, dim=128, levels=3, in_channels=3):
        if in_channels != 3:
            raise EwvitError("native MWT supports in_channels=3 (9 high-frequency planes) only")
        if dim % 128 != 0:
            raise EwvitError(f"native MWT needs dama_dim to be a multiple of 128 (got {dim})")
        if levels != 3:
            raise EwvitError("native MWT implements the reference's 3-level decomposition (model.py:35) only")
        self.dim, self.levels = dim, levels
        sd, target = _host_sd(sd)
        dev = torch.device("cpu")
        # hf_conv.seperate: three Conv2d(3,18)+BN, stacked
        self.head_w = torch.stack([sd[f"hf_conv.seperate.{i}.0.weight"].float() for i in range(3)]).contiguous()
        sc, sh = zip(*[_fold_bn(sd, f"hf_conv.seperate.{i}.0.", f"hf_conv.seperate.{i}.1.") for i in range(3)])
        self.head_scale, self.head_shift = torch.cat(sc).contiguous(), torch.cat(sh).contiguous()
        # tensor-core head: the three convs as one block-diagonal [64, 9 x 16] bf16 matrix, k = (dy*3 + dx)*16 + (3g + ic)
        wbd = torch.zeros((64, 3, 3, 16), dtype=torch.float32, device=dev)
        for g in range(3):
            wg = self.head_w[g]                                   # [18, 3, ky, kx]
            wbd[g * 18:(g + 1) * 18, :, :, g * 3:(g + 1) * 3] = wg.permute(0, 2, 3, 1)
        self.head_scale64 = torch.zeros(64, dtype=torch.float32, device=dev)
        self.head_shift64 = torch.zeros(64, dtype=torch.float32, device=dev)
        self.head_scale64[:54] = self.head_scale
        self.head_shift64[:54] = self.head_shift
        self.head_w3 = pack_head3_weights(wbd)
        self.head_scale192 = self.head_scale64.repeat(3).contiguous()      # the head is shared by the three levels
        self.head_shift192 = self.head_shift64.repeat(3).contiguous()
        self.fus_w = _conv_w_tapmajor(sd["hf_conv.fusion.0.weight"], 64)
        self.fus_scale, self.fus_shift = _fold_bn(sd, "hf_conv.fusion.0.", "hf_conv.fusion.1.")
        self.ms_w = _conv_w_tapmajor(sd["multiscale_fusion.0.weight"])
        self.ms_scale, self.ms_shift = _fold_bn(sd, "multiscale_fusion.0.", "multiscale_fusion.1.")
        self.fc_w = _conv_w_tapmajor(sd["freq_conv.0.weight"])
        self.fc_scale, self.fc_shift = _fold_bn(sd, "freq_conv.0.", "freq_conv.1.")
        self.fp_w = _conv_w_tapmajor(sd["freq_pool.1.weight"])
        self.fp_scale, self.fp_shift = _fold_bn(sd, "freq_pool.1.", "freq_pool.2.")
        self.__dict__.update(_to_device(self.__dict__, target))
        self.device = target
        self._ws = {}

    def _workspace(self, n, h, w):
        """One grow-only workspace per frame size (~22 MB per frame): allocated for the largest n seen, smaller batches
        (ragged tails of the macro-batch splitter) use leading slices.  Single-stream: two concurrent forwards of one
        runner on different CUDA streams would share these buffers."""
        full = self._ws.get((h, w))
        if full is None or full["n"] < n:
            self._ws.pop((h, w), None)
            full = None
            if len(self._ws) > 1:           # a third frame size: drop the others rather than hold several multi-GB sets
                self._ws.clear()
            dev, bf = self.device, torch.bfloat16
            h1, w1, d = h // 2, w // 2, self.dim
            h2, w2 = (h1 - 1) // 2 + 1, (w1 - 1) // 2 + 1          # freq_conv, stride 2
            h4, w4 = (h2 // 2 - 1) // 2 + 1, (w2 // 2 - 1) // 2 + 1  # maxpool then stride-2 conv
            full = {
                "n": n,
                "hf": [torch.empty((n, 3, 3, h >> l, w >> l), dtype=torch.float32, device=dev) for l in (1, 2, 3)],
                "up": torch.zeros((n, h1 + 2, w1 + 2, 32), dtype=bf, device=dev),        # channel 9 l + c per pixel; zero border, kept zero
                "head": torch.zeros((n, h1 + 2, w1 + 2, 192), dtype=bf, device=dev),     # [level][64] per pixel; zero border, kept zero
                "cat": torch.empty((n, h1 + 2, w1 + 2, 3 * d), dtype=bf, device=dev),
                "ms": torch.empty((n, h1 + 2, w1 + 2, d), dtype=bf, device=dev),
                "fc": torch.empty((n, h2, w2, d), dtype=bf, device=dev),
                "mp": torch.empty((n, h2 // 2, w2 // 2, d), dtype=bf, device=dev),
                "pc": torch.empty((n, h4, w4, d), dtype=bf, device=dev),
            }
            self._ws[(h, w)] = full
        if full["n"] == n:
            return full
        return {k: ([t[:n] for t in v] if isinstance(v, list) else v[:n]) for k, v in full.items() if k != "n"}

    def forward(self, frames, out=None, norm=None):
        """frames [n,3,H,W] fp32 CUDA (or uint8 with norm=(mean, std)), H, W multiples of 8 -> [n, dim] fp32."""
        n, c, h, w = frames.shape
        if c != 3 or h % 8 or w % 8:
            raise EwvitError(f"native MWT needs [n,3,H,W] with H,W multiples of 8 (got {tuple(frames.shape)})")
        ws = self._workspace(n, h, w)
        h1, w1, d = h // 2, w // 2, self.dim
        hf = ws["hf"]
        with stage("mwt.dwt3"):
            ops.dwt3_haar(frames, out={"hf1": hf[0], "hf2": hf[1], "hf3": hf[2]}, want=("hf1", "hf2", "hf3"), norm=norm)
        with stage("mwt.head"):       # all three levels: one upsample launch, one block-diagonal tensor-core conv
            ops.mwt_upsample3(*[hf[lvl].view(n, 9, h >> (lvl + 1), w >> (lvl + 1)) for lvl in range(3)], ws["up"], h1, w1)
            ops.mwt_head_conv3(ws["up"], self.head_w3, self.head_scale192, self.head_shift192, ws["head"], h1, w1)
        with stage("mwt.hf_fusion"):
            for lvl in range(3):        # level lvl = channels [64 lvl, 64 lvl + 64) of every head pixel
                ops.conv3x3_bf16(ws["head"], self.fus_w, n, h1, w1, 1, True, self.fus_scale, self.fus_shift, True,
                                 ws["cat"], lvl * d, True, x_coff=64 * lvl)
        with stage("mwt.multiscale"):
            ops.conv3x3_bf16(ws["cat"], self.ms_w, n, h1, w1, 1, True, self.ms_scale, self.ms_shift, True, ws["ms"], 0, True)
        with stage("mwt.freq_conv"):
            ops.conv3x3_bf16(ws["ms"], self.fc_w, n, h1, w1, 2, True, self.fc_scale, self.fc_shift, True, ws["fc"], 0, False)
        with stage("mwt.pool_tail"):
            ops.maxpool2x2(ws["fc"], ws["mp"])
            hp, wp = ws["mp"].shape[1:3]
            ops.conv3x3_bf16(ws["mp"], self.fp_w, n, hp, wp, 2, False, self.fp_scale, self.fp_shift, True, ws["pc"], 0, False)
            res = ops.gap(ws["pc"], out)
        return res


# ------------------------------------------------------------------------------------------ SFE
def make_backbone(net: nn.Module, device, v2s: bool):
    """Native runner of an EfficientNet feature extractor: torchvision V2-S ``features`` (the dynamic path, sfe.py:111-113)
    or the efficientnet_pytorch-style b0 module (the two ablation branches, sfe.py:109,148).  There is no library path."""
    return NativeEffNetV2(net, device) if v2s else NativeEffNetB0(net, device)


def _fold_conv_bn(conv, bn):
    """fp32 (weight, bias) of conv followed by eval-mode BatchNorm."""
    w = conv.weight.detach().float()
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv.bias is not None:
        bias = bias + conv.bias.detach().float() * scale
    return w * scale.view(-1, 1, 1, 1), bias.contiguous()


def _w3x3_tapmajor_padded(w):
    """[cout, cin, 3, 3] fp32 -> bf16 [cout, Kpad]: dense k = (ky*3+kx)*cin + c, zero-padded to a multiple of 64
    (layout of ewvit_conv_nhwc_bf16)."""
    cout, cin = w.shape[:2]
    k = 9 * cin
    kpad = (k + 63) // 64 * 64
    out = torch.zeros((cout, kpad), dtype=torch.bfloat16, device=w.device)
    out[:, :k] = w.permute(0, 2, 3, 1).reshape(cout, k).to(torch.bfloat16)
    return out.contiguous()


def _w3x3_window_packed(w):
    """[cout, cin, 3, 3] fp32 (cin < 64) -> bf16 [cout, 3*nsub*64] for the overlapping-window conv path:
    k = dy*(nsub*64) + dx*cin + c with nsub = ceil(3*cin/64), zero elsewhere (ewvit_conv_nhwc_bf16_ex)."""
    cout, cin = w.shape[:2]
    nsub = (3 * cin + 63) // 64
    out = torch.zeros((cout, 3, nsub * 64), dtype=torch.bfloat16, device=w.device)
    out[:, :, :3 * cin] = w.permute(0, 2, 3, 1).reshape(cout, 3, 3 * cin).to(torch.bfloat16)     # [cout, dy, (dx, c)]
    return out.reshape(cout, 3 * nsub * 64).contiguous()


class NativeEffNetV2:
    """torchvision ``efficientnet_v2_s(...).features`` (sfe.py:111-113,150) in eval mode on the native kernels:
    FusedMBConv 3x3 convs, 1x1 expand/project convs and the 1x1 head on the tcgen05 implicit-GEMM kernel with
    bias + SiLU + residual fused into the epilogue; depthwise 3x3 + SiLU + SE squeeze in one kernel; SE gate +
    scaling in one kernel; the stem reads the fp32 NCHW frames directly.  NHWC bf16 activations, BatchNorm folded."""

    def __init__(self, features: nn.Module, device):
        target = torch.device(device)
        dev = torch.device("cpu")           # fold and pack on the host, upload once at the end
        if any(p.device.type != "cpu" for p in features.parameters()):
            features = copy.deepcopy(features).to("cpu")
        self.ops = []
        self.win_w, self.cin3 = {}, {}      # op index -> window-packed weights / input channels of the 3x3 convs
        self._padbuf = {}
        mods = list(features)
        stem = mods[0]
        w, b = _fold_conv_bn(stem[0], stem[1])
        if tuple(stem[0].kernel_size) != (3, 3) or tuple(stem[0].stride) != (2, 2) or w.shape[1] != 3:
            raise EwvitError("native backbone: unexpected stem")
        self.ops.append(("stem", w.to(dev).contiguous(), b.to(dev)))
        for stage in mods[1:-1]:
            for blk in stage:
                kind = type(blk).__name__
                layers = list(blk.block)
                res = bool(blk.use_res_connect)
                if kind == "FusedMBConv":
                    c0 = layers[0]
                    w, b = _fold_conv_bn(c0[0], c0[1])
                    stride = c0[0].stride[0]
                    # candidate for the overlapping-window path (needs padded layouts).  Measured: 48 -> 192 @56^2 0.52 -> 0.36 ms,
                    # but 24 -> 24 @112^2 0.63 -> 0.85 ms (its 48-byte pixel pitch makes every 128-byte TMA row straddle two
                    # cache lines and K pads 72 -> 128 per tap row), so the 24-channel layers stay on the assembled-A path
                    if 32 <= w.shape[1] < 64 and stride == 1:
                        self.win_w[len(self.ops)] = _w3x3_window_packed(w).to(dev)
                    self.cin3[len(self.ops)] = int(w.shape[1])
                    if len(layers) == 1:
                        self.ops.append(("conv3", _w3x3_tapmajor_padded(w).to(dev), b.to(dev), stride, "silu", res, True))
                    else:
                        self.ops.append(("conv3", _w3x3_tapmajor_padded(w).to(dev), b.to(dev), stride, "silu", False, False))
                        w2, b2 = _fold_conv_bn(layers[1][0], layers[1][1])
                        self.ops.append(("conv1", w2.flatten(1).to(torch.bfloat16).to(dev).contiguous(), b2.to(dev), None, res, True))
                elif kind == "MBConv":
                    ex, dw, se, pr = layers
                    w, b = _fold_conv_bn(ex[0], ex[1])
                    self.ops.append(("conv1", w.flatten(1).to(torch.bfloat16).to(dev).contiguous(), b.to(dev), "silu", False, False))
                    wd_, bd = _fold_conv_bn(dw[0], dw[1])
                    c = wd_.shape[0]
                    self.ops.append(("dw", wd_.reshape(c, 9).t().contiguous().to(dev), bd.to(dev), dw[0].stride[0]))
                    self.ops.append(("se", se.fc1.weight.detach().float().flatten(1).contiguous().to(dev),
                                     se.fc1.bias.detach().float().to(dev),
                                     se.fc2.weight.detach().float().flatten(1).t().contiguous().to(dev),
                                     se.fc2.bias.detach().float().to(dev)))
                    w2, b2 = _fold_conv_bn(pr[0], pr[1])
                    # project conv: the SE gate is applied while its A operand is assembled ("conv1g")
                    self.ops.append(("conv1g", w2.flatten(1).to(torch.bfloat16).to(dev).contiguous(),
                                     b2.to(dev), None, res, True))
                else:
                    raise EwvitError(f"native backbone: unsupported block {kind}")
        head = mods[-1]
        w, b = _fold_conv_bn(head[0], head[1])
        self.ops.append(("conv1", w.flatten(1).to(torch.bfloat16).to(dev).contiguous(), b.to(dev), "silu", False, False))
        self.layout = self._plan_layouts(True)
        # SiLU epilogues evaluate h*tanh(h) + h with h = v/2: fold the 1/2 into weights and bias (exact: a power of two)
        for i, op in enumerate(self.ops):
            if op[0] == "conv3" and op[4] == "silu" and not self._is_c24(i):
                self.ops[i] = ("conv3", op[1] * 0.5, op[2] * 0.5, op[3], "silu_h", op[5], op[6])
                if i in self.win_w:
                    self.win_w[i] = self.win_w[i] * 0.5
            elif op[0] == "conv1" and op[3] == "silu":
                self.ops[i] = ("conv1", op[1] * 0.5, op[2] * 0.5, "silu_h", op[4], op[5])
        self.ops = _to_device(self.ops, target)
        self.win_w = _to_device(self.win_w, target)
        self.device = target

    def _is_c24(self, i):
        op = self.ops[i]
        return op[0] == "conv3" and self.cin3[i] == 24 and op[1].shape[0] == 24 and op[3] == 1 and op[4] == "silu"

    def _plan_layouts(self, enable):
        """Per op (in_padded, out_padded).  The stride-1 3x3 convs with cin < 64 (stages 1-2 of V2-S) run fastest on
        "padded-flat" tensors [n, h+2, w+2, c] (overlapping-window TMA path, no im2col); the 1x1 convs between them keep
        that layout, the stride-2 small-channel convs convert (they read the interior of a padded input and can write
        either layout).  Falls back to plain layouts everywhere if a padded tensor would reach an op that cannot take it."""
        plain = [(False, False)] * len(self.ops)
        if not enable:
            return plain

        def wants_padded(j):
            while j < len(self.ops) and self.ops[j][0] == "conv1":
                j += 1
            return j < len(self.ops) and self.ops[j][0] == "conv3" and j in self.win_w

        plan, padded = [], False
        for i, op in enumerate(self.ops):
            kind = op[0]
            if kind == "stem":
                plan.append((False, wants_padded(i + 1)))
            elif kind == "conv3":
                if i in self.win_w and padded:
                    plan.append((True, True))
                elif self.cin3[i] < 64:
                    plan.append((padded, wants_padded(i + 1)))
                elif padded:
                    return plain
                else:
                    plan.append((False, False))
            elif kind == "conv1":
                plan.append((padded, padded))
            else:
                if padded:
                    return plain
                plan.append((False, False))
            padded = plan[-1][1]
        return plain if padded else plan

    def _zero_bordered(self, i, shape):
        """Persistent output buffer of op i whose border is zeroed once (the op rewrites the interior every call).
        Grow-only in the frame count; smaller batches use a leading slice."""
        key = (i, tuple(shape[1:]))
        t = self._padbuf.get(key)
        if t is None or t.shape[0] < shape[0]:
            if len(self._padbuf) > 16:
                self._padbuf.clear()
            t = torch.zeros(shape, dtype=torch.bfloat16, device=self.device)
            self._padbuf[key] = t
        return t[: shape[0]]

    def forward(self, frames, norm=None):
        """fp32 [n,3,H,W] (or uint8 with norm=(mean, std)) -> bf16 NHWC [n, H/32, W/32, C_out]."""
        block_in = None      # input of the current residual block
        x = None
        pooled = None
        gate = None
        for i, op in enumerate(self.ops):
            kind = op[0]
            with stage(f"bb.{kind}" if TIMER is None or not getattr(TIMER, "per_layer", False) else f"bb.{i:03d}.{kind}"):
                in_p, out_p = self.layout[i]
                if kind == "stem":
                    if out_p:
                        n, _, h, wd = frames.shape
                        buf = self._zero_bordered(i, (n, (h - 1) // 2 + 3, (wd - 1) // 2 + 3, op[1].shape[0]))
                        x = ops.stem_conv(frames, op[1], op[2], out=buf, out_padded=True, norm=norm)
                    else:
                        x = ops.stem_conv(frames, op[1], op[2], norm=norm)
                    block_in = x
                elif kind == "conv3":
                    _, w, b, stride, act, res, ends = op
                    if in_p or out_p:
                        window = in_p and out_p and i in self.win_w
                        out = None
                        if out_p and not window:       # stride-2 conv into a padded layout: interior only
                            n, hp, wp = x.shape[0], x.shape[1] - 2 * in_p, x.shape[2] - 2 * in_p
                            out = self._zero_bordered(i, (n, (hp - 1) // stride + 3, (wp - 1) // stride + 3, w.shape[0]))
                        x = ops.conv_nhwc_bf16_ex(x, self.win_w[i] if window else w, 3, stride, self.cin3[i], bias=b, act=act,
                                                  residual=block_in if res else None, out=out, in_padded=in_p, out_padded=out_p)
                    elif self._is_c24(i) and (not res or block_in is x):
                        x = ops.conv3x3_c24(x, w, b, residual=res)       # direct conv on warp-level MMAs (N = 24)
                    else:
                        x = ops.conv_nhwc_bf16(x, w, 3, stride, bias=b, act=act, residual=block_in if res else None)
                    if ends:
                        block_in = x
                elif kind == "conv1":
                    _, w, b, act, res, ends = op
                    if in_p:
                        x = ops.conv_nhwc_bf16_ex(x, w, 1, 1, w.shape[1], bias=b, act=act, residual=block_in if res else None,
                                                  in_padded=True, out_padded=True)
                    else:
                        x = ops.conv_nhwc_bf16(x, w, 1, 1, bias=b, act=act, residual=block_in if res else None)
                    if ends:
                        block_in = x
                elif kind == "dw":
                    _, w, b, stride = op
                    n, h, wd, c = x.shape
                    pooled = torch.empty((n, c), dtype=torch.float32, device=x.device)
                    x = ops.dwconv3x3(x, w, b, stride, pooled=pooled)
                elif kind == "se":
                    gate = ops.se_gate(pooled, op[1], op[2], op[3], op[4], bf16=True)      # applied inside the project conv's operand path
                elif kind == "conv1g":
                    _, w, b, act, res, ends = op
                    x = ops.conv1x1_gated(x, gate, w, bias=b, act=act, residual=block_in if res else None)
                    if ends:
                        block_in = x
        return x


class NativeEffNetB0:
    """``EfficientNet.from_pretrained('efficientnet-b0').extract_features`` (efficientnet_pytorch layout: ``_conv_stem``, ``_bn0``,
    ``_blocks[i]._expand_conv/_bn0/_depthwise_conv/_bn1/_se_reduce/_se_expand/_project_conv/_bn2``, ``_conv_head``, ``_bn1``;
    network/sfe.py:109,148) in eval mode on the native kernels: TF-'SAME' stem, 1x1 expand / SE-gated project / head convs on the
    tcgen05 GEMM kernel, 3x3 / 5x5 depthwise + SiLU + squeeze in ``ewvit_dwconv_nhwc_bf16``, SE gate kernel.  NHWC bf16, BN folded
    (eps 1e-3 as the module says).  Geometry is read off the module's own convolutions, so the real package and the built-in
    stand-in (``network/_effnet_b0.py``) are handled alike."""

    def __init__(self, net: nn.Module, device):
        target = torch.device(device)
        if any(p.device.type != "cpu" for p in net.parameters()):
            net = copy.deepcopy(net).to("cpu")
        bf = torch.bfloat16
        one = lambda v: int(v[0] if isinstance(v, (tuple, list)) else v)
        w, b = _fold_conv_bn(net._conv_stem, net._bn0)
        if tuple(w.shape[1:]) != (3, 3, 3) or one(net._conv_stem.stride) != 2 or w.shape[0] > 32:
            raise EwvitError("native b0: unexpected stem")
        self.ops = [("stem", w.contiguous(), b)]
        for blk in net._blocks:
            dwc = blk._depthwise_conv
            mid, k, stride = dwc.weight.shape[0], one(dwc.kernel_size), one(dwc.stride)
            cin = blk._expand_conv.weight.shape[1] if hasattr(blk, "_expand_conv") else mid
            cout = blk._project_conv.weight.shape[0]
            if hasattr(blk, "_expand_conv"):
                w, b = _fold_conv_bn(blk._expand_conv, blk._bn0)
                self.ops.append(("conv1", (w.flatten(1) * 0.5).to(bf).contiguous(), b * 0.5, "silu_h", False, False))   # halved: h*tanh(h)+h
            w, b = _fold_conv_bn(dwc, blk._bn1)
            self.ops.append(("dw", w.reshape(mid, k * k).t().contiguous(), b, k, stride))
            self.ops.append(("se", blk._se_reduce.weight.detach().float().flatten(1).contiguous(), blk._se_reduce.bias.detach().float(),
                             blk._se_expand.weight.detach().float().flatten(1).t().contiguous(), blk._se_expand.bias.detach().float()))
            w, b = _fold_conv_bn(blk._project_conv, blk._bn2)
            self.ops.append(("conv1g", w.flatten(1).to(bf).contiguous(), b, None, stride == 1 and cin == cout, True))
        w, b = _fold_conv_bn(net._conv_head, net._bn1)
        self.ops.append(("conv1", (w.flatten(1) * 0.5).to(bf).contiguous(), b * 0.5, "silu_h", False, False))
        self.ops = _to_device(self.ops, target)
        self.device = target

    def forward(self, frames, norm=None):
        """fp32 [n,3,H,W] -> bf16 NHWC [n, H/32, W/32, 1280]."""
        if norm is not None or frames.dtype != torch.float32:
            raise EwvitError("native b0 takes normalised fp32 frames")
        x = block_in = pooled = gate = None
        for i, op in enumerate(self.ops):
            kind = op[0]
            with stage(f"b0.{kind}"):
                if kind == "stem":
                    x = block_in = ops.stem_conv(frames, op[1], op[2], same_tf=True)
                elif kind == "conv1":
                    x = ops.conv_nhwc_bf16(x, op[1], 1, 1, bias=op[2], act=op[3])
                elif kind == "dw":
                    x, pooled = ops.dwconv(x, op[1], op[2], op[3], op[4], same_tf=True, pooled=True)
                elif kind == "se":
                    gate = ops.se_gate(pooled, op[1], op[2], op[3], op[4], bf16=True)
                else:       # conv1g: SE gate applied in the operand path, + skip connection
                    x = block_in = ops.conv1x1_gated(x, gate, op[1], bias=op[2], act=None, residual=block_in if op[4] else None)
        return x


class SfeRunner:
    """Native ``EfficientViT.forward`` after the backbone; ``backbone`` maps fp32 frames to the bf16 NHWC
    feature map.  ``sd`` holds the module's own keys (``patch_to_embedding.weight`` ...)."""

    def __init__(self, sd, cfg, backbone, output_mode="feature_map"):
        m = cfg["model"]
        self.dim, self.depth, self.heads, self.dim_head = m["dim"], m["depth"], m["heads"], m["dim-head"]
        self.mlp_dim, self.emb_dim = m["mlp-dim"], m["emb-dim"]
        self.output_mode = output_mode
        sd, target = _host_sd(sd)
        bf = torch.bfloat16
        f32 = lambda k: sd[k].float().contiguous()
        self.patch_w = sd["patch_to_embedding.weight"].to(bf).contiguous()
        self.patch_b = f32("patch_to_embedding.bias")
        self.cls = f32("cls_token").reshape(-1)
        self.pos = f32("pos_embedding").reshape(-1, self.dim)
        self.layers = []
        for l in range(self.depth):
            q = f"transformer.layers.{l}."
            self.layers.append({
                "ln1": (f32(q + "0.norm.weight"), f32(q + "0.norm.bias")),
                "qkv": sd[q + "0.fn.to_qkv.weight"].to(bf).contiguous(),
                "out_w": sd[q + "0.fn.to_out.0.weight"].to(bf).contiguous(), "out_b": f32(q + "0.fn.to_out.0.bias"),
                "ln2": (f32(q + "1.norm.weight"), f32(q + "1.norm.bias")),
                "ff1_w": sd[q + "1.fn.net.0.weight"].to(bf).contiguous(), "ff1_b": f32(q + "1.fn.net.0.bias"),
                "ff2_w": sd[q + "1.fn.net.3.weight"].to(bf).contiguous(), "ff2_b": f32(q + "1.fn.net.3.bias"),
            })
        for w in (self.patch_w, *[l[k] for l in self.layers for k in ("qkv", "out_w", "ff1_w", "ff2_w")]):
            if w.shape[0] % 128 or w.shape[1] % 64:
                raise EwvitError("native SFE needs every Linear to have out %128 == 0 and in %64 == 0 "
                                 f"(architecture.yaml gives {tuple(w.shape)})")
        if output_mode == "cls":
            self.h1_w = sd["mlp_head.0.weight"].to(bf).contiguous()
            self.h1_b = f32("mlp_head.0.bias")
            w2 = sd["mlp_head.2.weight"]
            self.num_classes = w2.shape[0]
            self.h2_w = torch.zeros((128, w2.shape[1]), dtype=bf, device=w2.device)
            self.h2_w[: w2.shape[0]] = w2.to(bf)
            self.h2_b = torch.zeros(128, dtype=torch.float32, device=w2.device)
            self.h2_b[: w2.shape[0]] = sd["mlp_head.2.bias"].float()
        else:
            fw = sd["feat_map.0.weight"]
            self.feat_dim = fw.shape[0]
            pad = (-fw.shape[0]) % 128
            self.fm_w = torch.cat([fw.to(bf), torch.zeros((pad, fw.shape[1]), dtype=bf, device=fw.device)]).contiguous()
            self.fm_b = torch.cat([sd["feat_map.0.bias"].float(), torch.zeros(pad, device=fw.device)]).contiguous()
        self.__dict__.update(_to_device(self.__dict__, target))
        self.backbone = backbone
        self._ws = {}

    def _workspace(self, n, dev):
        """Grow-only token workspace (largest n seen; smaller batches use leading slices)."""
        inner = self.heads * self.dim_head
        kblocks = self.patch_w.shape[1] // 64
        splits = max(1, min(kblocks, 148 // max(1, math.ceil(n / 128) * (self.dim // 128))))
        full = self._ws
        if not full or full["n"] < n or full["dev"] != dev:
            f32, bf = torch.float32, torch.bfloat16
            cap = max(n, 64)
            max_splits = max(1, min(kblocks, 148 // max(1, self.dim // 128)))
            full = self._ws = {
                "n": cap, "dev": dev,
                "ws": torch.empty((max_splits * cap * self.dim,), dtype=f32, device=dev),
                "emb": torch.empty((cap, self.dim), dtype=f32, device=dev),
                "x": [torch.empty((2 * cap, self.dim), dtype=f32, device=dev) for _ in range(2)],
                "xn": torch.empty((2 * cap, self.dim), dtype=bf, device=dev),
                "qkv": torch.empty((2 * cap, 3 * inner), dtype=f32, device=dev),
                "att": torch.empty((2 * cap, inner), dtype=bf, device=dev),
                "hid": torch.empty((2 * cap, self.mlp_dim), dtype=bf, device=dev),
                "tok": torch.empty((cap, self.dim), dtype=bf, device=dev),
            }
        return {
            "splits": splits,
            "ws": full["ws"][: splits * n * self.dim].view(splits, n, self.dim) if splits > 1 else None,
            "emb": full["emb"][:n], "x": [t[: 2 * n] for t in full["x"]], "xn": full["xn"][: 2 * n],
            "qkv": full["qkv"][: 2 * n], "att": full["att"][: 2 * n], "hid": full["hid"][: 2 * n], "tok": full["tok"][:n],
        }

    def head(self, feat_nhwc, pos_index, out=None):
        """feat_nhwc: bf16 [n, ph*pw*channels] (NHWC flatten of the backbone map); pos_index int32 [n]."""
        n = feat_nhwc.shape[0]
        ws = self._workspace(n, feat_nhwc.device)
        with stage("sfe.patch_embed"):
            ops.linear_bf16(feat_nhwc, self.patch_w, shift=self.patch_b, out=ws["emb"], splits=ws["splits"], workspace=ws["ws"])
        with stage("sfe.vit"):
            return self._vit(ws, n, pos_index, out)

    def _vit(self, ws, n, pos_index, out):
        x, y = ws["x"]
        ops.vit_assemble(ws["emb"], self.cls, self.pos, pos_index, out=x)
        for L in self.layers:
            ops.layernorm_bf16(x, *L["ln1"], eps=LN_EPS, out=ws["xn"])
            ops.linear_bf16(ws["xn"], L["qkv"], out=ws["qkv"])
            ops.vit_attention(ws["qkv"], n, 2, self.heads, self.dim_head, out=ws["att"])
            ops.linear_bf16(ws["att"], L["out_w"], shift=L["out_b"], residual=x, out=y)
            ops.layernorm_bf16(y, *L["ln2"], eps=LN_EPS, out=ws["xn"])
            ops.linear_bf16(ws["xn"], L["ff1_w"], shift=L["ff1_b"], act="gelu", out=ws["hid"])
            ops.linear_bf16(ws["hid"], L["ff2_w"], shift=L["ff2_b"], residual=y, out=x)
        d = self.dim
        if self.output_mode == "cls":      # token 0 -> mlp_head (sfe.py:163-166)
            ops.layernorm_bf16(x, None, None, out=ws["tok"], rows=n, ldx=2 * d, d=d)
            hid = ops.linear_bf16(ws["tok"], self.h1_w, shift=self.h1_b, act="relu", out_dtype=torch.bfloat16)
            res = ops.linear_bf16(hid, self.h2_w, shift=self.h2_b)
            return res[:, : self.num_classes].contiguous()
        # token 1 -> feat_map Linear + ReLU (sfe.py:168-173)
        ops.layernorm_bf16(x[1], None, None, out=ws["tok"], rows=n, ldx=2 * d, d=d)   # rows 1, 3, 5, ...
        res = ops.linear_bf16(ws["tok"], self.fm_w, shift=self.fm_b, act="relu", out=out if self.fm_w.shape[0] == self.feat_dim else None)
        return res if res.shape[1] == self.feat_dim else res[:, : self.feat_dim].contiguous()

    def features(self, frames, norm=None):
        """fp32 frames [n,3,H,W] (uint8 with norm on the V2-S backbone) -> bf16 [n, 62720] NHWC-flattened backbone features
        (the reference's ``rearrange 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'`` at sfe.py:153 is this flatten for one patch)."""
        f = self.backbone.forward(frames, norm=norm)             # NHWC bf16
        return f.reshape(f.shape[0], -1)

    def forward(self, frames, pos_index, out=None, norm=None):
        with stage("sfe.backbone"):
            feat = self.features(frames, norm=norm)
            if not feat.is_contiguous():
                feat = feat.contiguous()
        return self.head(feat, pos_index, out)


# ------------------------------------------------------------------------------------------ DAMA
def pack_dama_weights(sd, dim, depth=2):
    """Flatten the cross-attention / gate weights into the layout of ``ewvit_dama_tail_fwd`` (include/ewvit.h)."""
    parts = []
    sd, target = _host_sd({k: v for k, v in sd.items() if k.startswith(("cross_att.", "fusion_gate.", "gate_net."))})
    f = lambda k: sd[k].float()
    for l in range(depth):
        for ln, att in ((0, 1), (2, 3)):
            q = f"cross_att.layers.{l}."
            parts += [f(f"{q}{ln}.weight"), f(f"{q}{ln}.bias"), f(f"{q}{att}.to_q.weight").t(), f(f"{q}{att}.to_kv.weight").t(),
                      f(f"{q}{att}.to_out.0.weight").t(), f(f"{q}{att}.to_out.0.bias")]
    centre = f("fusion_gate.0.weight")[:, :, 1, 1]              # [dim, 2*dim]: only the centre tap is live on a 1x1 map
    scale, shift = _fold_bn(sd, "fusion_gate.0.", "fusion_gate.1.")
    parts += [centre.t(), scale, shift, f("gate_net.2.weight").t(), f("gate_net.2.bias"), f("gate_net.5.weight"), f("gate_net.5.bias")]
    pack = torch.cat([p.contiguous().reshape(-1) for p in parts]).contiguous()
    assert pack.numel() == ops.dama_wpack_floats(dim, depth)
    return pack.to(target)


def chunk_pos_index(b, k, batch_size):
    """Position of every frame inside its reference chunk: ``DAMA.forward`` feeds ``x[:, s:e].flatten(0,1)``
    (dama.py:179-186) and ``EfficientViT.forward`` adds ``pos_embedding[0:N]`` along that flattened axis
    (sfe.py:158-159), so frame (b, k) of a chunk of length L gets row ``b*L + (k - s)``.  Raises like the
    reference's broadcast when a chunk holds more than ``emb-dim`` frames."""
    idx = torch.empty((b, k), dtype=torch.int32)
    for s in range(0, k, batch_size):
        e = min(s + batch_size, k)
        length = e - s
        idx[:, s:e] = torch.arange(b, dtype=torch.int32).view(b, 1) * length + torch.arange(length, dtype=torch.int32).view(1, length)
    return idx.reshape(-1)


def check_chunk_limit(b, k, batch_size, emb_dim):
    n = b * min(batch_size, k)
    if n > emb_dim:
        raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b ({emb_dim}) at "
                           "non-singleton dimension 0")


class DamaRunner:
    """Native ``DAMA.forward`` (dynamic mode).  ``sd`` holds the DAMA module's own keys."""

    def __init__(self, sd, cfg, backbone, dim=128, heads=4, levels=3, depth=2):
        self.dim, self.heads, self.depth = dim, heads, depth
        self.cfg = cfg
        self.sfe = SfeRunner(_sub(sd, "sfe."), cfg, backbone)
        self.mwt = MwtRunner(_sub(sd, "mwt."), dim=dim, levels=levels)
        self.wpack = pack_dama_weights(sd, dim, depth)
        self._pos_cache = {}

    def pos_index(self, b, k, batch_size, device):
        key = (b, k, batch_size, str(device))
        t = self._pos_cache.get(key)
        if t is None:
            if len(self._pos_cache) > 16:
                self._pos_cache.clear()
            t = chunk_pos_index(b, k, batch_size).to(device)
            self._pos_cache[key] = t
        return t

    def process_frames(self, frames, pos_index, norm=None):
        """frames [n,3,H,W] -> (fused, space, freq) each [n, dim] fp32 (``_process_frame``, dama.py:130-169)."""
        space = self.sfe.forward(frames, pos_index, norm=norm)
        freq = self.mwt.forward(frames, norm=norm)
        with stage("dama.tail"):
            return ops.dama_tail(space, freq, self.wpack, self.heads, self.depth, LN_EPS)

    def forward_frames(self, x, batch_size, norm=None):
        """x [B,K,3,H,W] fp32 CUDA (or uint8 with norm=(mean, std)) -> per-frame (fused, space, freq) [B*K, dim] in (b, k) order."""
        b, k = x.shape[:2]
        check_chunk_limit(b, k, batch_size, self.sfe.pos.shape[0])      # rows of the checkpoint's pos_embedding, not the yaml's
        frames = x.reshape(b * k, *x.shape[2:])
        if not frames.is_contiguous():
            frames = frames.contiguous()
        pos = self.pos_index(b, k, batch_size, x.device)
        n = b * k
        if n <= MACRO_BATCH:
            return self.process_frames(frames, pos, norm=norm)
        outs = [torch.empty((n, self.dim), dtype=torch.float32, device=x.device) for _ in range(3)]
        for s in range(0, n, MACRO_BATCH):
            e = min(s + MACRO_BATCH, n)
            part = self.process_frames(frames[s:e], pos[s:e], norm=norm)
            for o, p in zip(outs, part):
                o[s:e].copy_(p)
        return tuple(outs)


class DetectorRunner:
    """Native ``DeepfakeDetector.forward(x, batch_size, 'dynamic')`` (model.py:83-99)."""

    def __init__(self, sd, cfg, backbone, dim=128):
        self.dama = DamaRunner(_sub(sd, "dama."), cfg, backbone, dim=dim)
        f = lambda k: sd[k].detach().float().contiguous()
        self.classifier = (f("classifier.0.weight"), f("classifier.0.bias"), f("classifier.3.weight").reshape(-1).contiguous(),
                           f("classifier.3.bias"))

    def forward(self, x, batch_size, norm=None):
        b, k = x.shape[:2]
        fused, space, freq = self.dama.forward_frames(x, batch_size, norm=norm)
        mf, ms, mq, logits = ops.video_head(fused, space, freq, b, k, self.classifier)
        return {"logits": logits, "fused": mf, "space": ms, "freq": mq}
