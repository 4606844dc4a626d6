"""``network.dama`` -- drop-in for the reference's adaptive fusion module (network/dama.py:14-206).

``CrossAttention``, ``BidirectionalCrossTransformer`` and ``DAMA`` keep their constructor arguments, attribute
names (``sfe``, ``mwt``, ``gate_net``, ``cross_att``, ``fusion_gate``), ``_process_frame`` / ``forward``
signatures, returned dict keys and state_dict layout.  Eval-mode CUDA calls run the whole frame batch through
the native pipeline in one pass (the reference's serial chunk loop only survives as the position-embedding
index rule); training takes the PyTorch composition.
"""
import torch
from torch import nn

from ._native import NativeMixin, load_architecture_config
from .mwt import MWT
from .sfe import EfficientViT


class CrossAttention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, 2 * inner, bias=False)
        needs_projection = not (heads == 1 and dim_head == dim)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if needs_projection else nn.Identity()

    def forward(self, x, context=None, kv_include_self=False):
        b, t, _ = x.shape
        context = x if context is None else context
        if kv_include_self:
            context = torch.cat((x, context), dim=1)
        split = lambda z: z.reshape(b, z.shape[1], self.heads, -1).transpose(1, 2)
        q = split(self.to_q(x))
        k, v = (split(z) for z in self.to_kv(context).chunk(2, dim=-1))
        attn = self.attend(torch.matmul(q, k.transpose(-1, -2)) * self.scale)
        out = torch.matmul(attn, v).transpose(1, 2).reshape(b, t, -1)
        return self.to_out(out)


class BidirectionalCrossTransformer(nn.Module):
    def __init__(self, dim, depth=1, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([nn.LayerNorm(dim), CrossAttention(dim, heads=heads, dim_head=dim_head, dropout=dropout),
                           nn.LayerNorm(dim), CrossAttention(dim, heads=heads, dim_head=dim_head, dropout=dropout)])
            for _ in range(depth)])

    def forward(self, space_tokens, freq_tokens):
        for space_norm, space_from_freq, freq_norm, freq_from_space in self.layers:
            space_tokens = space_tokens + space_from_freq(space_norm(space_tokens), freq_tokens, kv_include_self=True)
            freq_tokens = freq_tokens + freq_from_space(freq_norm(freq_tokens), space_tokens, kv_include_self=True)
        return space_tokens, freq_tokens


class DAMA(NativeMixin, nn.Module):
    def __init__(self, in_channels=3, dim=128, num_heads=4, levels=3, batch_size=16):
        super().__init__()
        self.dim = dim
        self.levels = levels
        self.batch_size = batch_size
        self.num_heads = num_heads
        self._config = load_architecture_config()
        self.sfe = EfficientViT(config=self._config, channels=1280, selected_efficient_net=1, feat_dim=dim,
                                output_mode="feature_map")
        self.mwt = MWT(in_channels=in_channels, dama_dim=dim, levels=levels)
        self.gate_net = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(2 * dim, dim // 2), nn.ReLU(),
                                      nn.Dropout(0.1), nn.Linear(dim // 2, 3), nn.Softmax(dim=1))
        self.cross_att = BidirectionalCrossTransformer(dim=dim, depth=2, heads=num_heads, dim_head=dim // num_heads,
                                                       dropout=0.1)
        self.fusion_gate = nn.Sequential(nn.Conv2d(2 * dim, dim, kernel_size=3, padding=1), nn.BatchNorm2d(dim),
                                         nn.ReLU(inplace=True))

    # ---- PyTorch composition (training / hooks); mirrors dama.py:130-169
    def _process_frame_torch(self, frame):
        b = frame.shape[0]
        space, freq = self.sfe(frame), self.mwt(frame)
        h_out, w_out = space.shape[-2:]
        s, f = self.cross_att(space.flatten(2).transpose(1, 2), freq.flatten(2).transpose(1, 2))
        space = s.transpose(1, 2).reshape(b, self.dim, h_out, w_out)
        freq = f.transpose(1, 2).reshape(b, self.dim, h_out, w_out)
        both = torch.cat([space, freq], dim=1)
        fused = self.fusion_gate(both)
        g = self.gate_net(both)
        mix = g[:, 0].view(b, 1, 1, 1) * space + g[:, 1].view(b, 1, 1, 1) * freq + g[:, 2].view(b, 1, 1, 1) * fused
        return {"fused": mix.mean(dim=[2, 3]), "space": space.mean(dim=[2, 3]), "freq": freq.mean(dim=[2, 3])}

    def _build_runner(self):
        from ewvit.engine import DamaRunner, make_backbone
        backbone = make_backbone(self.sfe.efficient_net.features, self.sfe.pos_embedding.device, v2s=True)
        sd = {k: v for k, v in self.state_dict().items() if not k.startswith("sfe.efficient_net.")}
        return DamaRunner(sd, self._config, backbone, dim=self.dim, heads=self.num_heads, levels=self.levels, depth=2)

    def _process_frame(self, frame):
        if self._use_native(frame):
            from ewvit.engine import check_chunk_limit
            n = frame.shape[0]
            check_chunk_limit(n, 1, 1, self.sfe.pos_embedding.shape[0])
            runner = self._native_runner(self._build_runner)
            pos = torch.arange(n, dtype=torch.int32, device=frame.device)
            fused, space, freq = runner.process_frames(frame.float().contiguous(), pos)
            return {"fused": fused, "space": space, "freq": freq}
        return self._process_frame_torch(frame)

    def forward(self, x, batch_size=16):
        """x [B, K, C, H, W] -> per-video means {'fused','space','freq'} [B, dim] (dama.py:171-206)."""
        b, k = x.shape[:2]
        if self._use_native(x):
            from ewvit import ops
            runner = self._native_runner(self._build_runner)
            fused, space, freq = runner.forward_frames(x.float(), batch_size)
            mf, ms, mq, _ = ops.video_head(fused, space, freq, b, k, None)
            return {"fused": mf, "space": ms, "freq": mq}
        acc = {name: torch.zeros(b, self.dim, device=x.device) for name in ("fused", "space", "freq")}
        for start in range(0, k, batch_size):
            feats = self._process_frame(x[:, start:min(start + batch_size, k)].flatten(0, 1))
            for name in acc:
                acc[name] = acc[name] + feats[name].view(b, -1, self.dim).sum(dim=1)
        return {name: v / k for name, v in acc.items()}
