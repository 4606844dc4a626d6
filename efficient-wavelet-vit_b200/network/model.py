"""``network.model.DeepfakeDetector`` -- drop-in for the reference's top-level detector
(network/model.py:9-171): same constructor, sub-module names (``dama``, ``mwt``, ``sfe``, ``sfe_cls``,
``fusion_gate``, ``feat_pooler``, ``classifier``), ``forward(x, batch_size, ablation)`` signature, returned
dict keys, ``configure_ablation`` and state_dict layout.
"""
import torch
from torch import nn
from torch.nn import functional as F

from ._native import NativeMixin, load_architecture_config
from .dama import DAMA
from .mwt import MWT
from .sfe import EfficientViT


class DeepfakeDetector(NativeMixin, nn.Module):
    def __init__(self, in_channels=3, dama_dim=128, batch_size=16, ablation="dynamic"):
        super().__init__()
        self.dama_dim = dama_dim
        self.in_channels = in_channels
        self.batch_size = batch_size
        self.ablation_config = ["dynamic", "sfe_only", "sfe_mwt"]
        self.config = load_architecture_config()

        self.dama = DAMA(in_channels=in_channels, dim=dama_dim, num_heads=4, levels=3, batch_size=batch_size)
        self.mwt = MWT(in_channels=in_channels, dama_dim=dama_dim)
        self.sfe = EfficientViT(config=self.config, channels=1280, feat_dim=dama_dim, selected_efficient_net=0)
        self.sfe_cls = EfficientViT(config=self.config, channels=1280, feat_dim=dama_dim, selected_efficient_net=0,
                                    output_mode="cls")
        self.fusion_gate = nn.Sequential(nn.Linear(2 * dama_dim, 2), nn.ReLU(), nn.Dropout(0.1))
        self.feat_pooler = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential(nn.Linear(dama_dim, 64), nn.ReLU(), nn.Dropout(0.3), nn.Linear(64, 1))

    def _chunks(self, x):
        k = x.shape[1]
        for start in range(0, k, self.batch_size):
            yield x[:, start:min(start + self.batch_size, k)].flatten(0, 1)

    def _build_runner(self):
        from ewvit.engine import DetectorRunner, make_backbone
        backbone = make_backbone(self.dama.sfe.efficient_net.features, self.classifier[0].weight.device, v2s=True)
        sd = {k: v for k, v in self.state_dict().items()
              if (k.startswith("dama.") and not k.startswith("dama.sfe.efficient_net.")) or k.startswith("classifier.")}
        return DetectorRunner(sd, self.config, backbone, dim=self.dama_dim)

    def _native_tensors(self):
        # only the tensors the dynamic path reads decide when its runner is rebuilt
        return [(k, t) for k, t in self.state_dict(keep_vars=True).items() if k.startswith("dama.") or k.startswith("classifier.")]

    def forward(self, x, batch_size, ablation):
        if batch_size is not None:
            self.batch_size = batch_size
        if ablation is not None:
            self.ablation = ablation
        b, k = x.shape[:2]

        if self.ablation == "dynamic":
            if self._use_native(x):
                return self._native_runner(self._build_runner).forward(x.float(), self.batch_size)
            feats = self.dama(x, batch_size=self.batch_size)
            return {"logits": self.classifier(feats["fused"]), "fused": feats["fused"], "space": feats["space"],
                    "freq": feats["freq"]}

        if self.ablation == "sfe_only":
            per_frame = torch.cat([self.sfe_cls(chunk).view(b, -1, 1) for chunk in self._chunks(x)], dim=1)
            return {"logits": per_frame.mean(dim=1), "model": "sfe_only"}

        if self.ablation == "sfe_mwt":
            sfe_parts, mwt_parts = [], []
            for chunk in self._chunks(x):
                sfe_parts.append(self.feat_pooler(self.sfe(chunk)).flatten(1).view(b, -1, self.dama_dim))
                mwt_parts.append(self.mwt(chunk).flatten(1).view(b, -1, self.dama_dim))
            sfe_mean = torch.cat(sfe_parts, dim=1).mean(dim=1)
            mwt_mean = torch.cat(mwt_parts, dim=1).mean(dim=1)
            gate = F.softmax(self.fusion_gate(torch.cat([sfe_mean, mwt_mean], dim=1)), dim=1)
            fused = sfe_mean * gate[:, 0:1] + mwt_mean * gate[:, 1:2]
            return {"logits": self.classifier(fused), "sfe": sfe_mean, "mwt": mwt_mean, "model": "sfe_mwt"}

    def forward_uint8(self, x, batch_size=None, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
        """Extension, not part of the reference API: score raw uint8 frames x [B,K,3,H,W].
        The reference's input pipeline turns decoded frames into floats with ``ToTensor`` and ``Normalize(mean, std)``
        (config/transforms.py:97-98) on the host; here that arithmetic happens on load inside the first two kernels (fused
        DWT, backbone stem), so a quarter of the bytes cross PCIe and HBM.  Bit-identical to
        ``forward(((x.float() / 255) - mean) / std, batch_size, 'dynamic')`` under ``torch.no_grad()``.  Eval mode, CUDA, dynamic
        ablation only; inference only (the outputs never carry a ``grad_fn``); ``self.ablation`` is left alone."""
        from ewvit import EwvitError
        if x.dtype != torch.uint8 or x.dim() != 5 or x.shape[2] != 3:
            raise EwvitError("forward_uint8: x must be a uint8 tensor [B, K, 3, H, W]")
        if batch_size is not None:
            self.batch_size = batch_size
        if self.training or not x.is_cuda:
            raise EwvitError("forward_uint8 serves eval-mode CUDA calls only (there is no CPU path)")
        key = (str(x.device), tuple(float(v) for v in mean), tuple(float(v) for v in std))
        cache = self.__dict__.setdefault("_ewvit_norm", {})
        norm = cache.get(key)
        if norm is None:        # built once: a host -> device copy per call would serialise the caller's copy/compute overlap
            norm = (torch.tensor(key[1], dtype=torch.float32, device=x.device), torch.tensor(key[2], dtype=torch.float32, device=x.device))
            cache[key] = norm
        return self._native_runner(self._build_runner).forward(x.contiguous(), self.batch_size, norm=norm)

    def configure_ablation(self, ablation):
        if ablation in self.ablation_config:
            self.ablation = ablation
        else:
            raise ValueError(f"Invalid ablation config: {ablation}.")
