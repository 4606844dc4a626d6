"""Shared plumbing of the drop-in modules: config loading, native-runner caching, dispatch rules."""
import os

import torch
import yaml

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_architecture_config():
    """The reference opens ``config/architecture.yaml`` relative to the CWD (dama.py:94, model.py:31);
    keep that, and fall back to the copy shipped with this package when the CWD has none."""
    for path in ("config/architecture.yaml", os.path.join(_PKG_ROOT, "config", "architecture.yaml")):
        if os.path.exists(path):
            with open(path, "r") as f:
                return yaml.safe_load(f)
    raise FileNotFoundError("config/architecture.yaml")


def force_torch_path():
    return os.environ.get("EWVIT_FORCE_TORCH", "0") == "1"


class NativeMixin:
    """Caches a native runner per module; rebuilt whenever a parameter/buffer was modified in place,
    replaced, or moved (sum of tensor versions + identities + device)."""

    def _native_signature(self):
        sig = 0
        dev = None
        for t in self.state_dict(keep_vars=True).values():
            sig += t._version + (id(t) & 0xFFFF)
            dev = t.device
        return sig, str(dev)

    def _native_runner(self, build):
        sig = self._native_signature()
        cache = self.__dict__.get("_ewvit_cache")
        if cache is None or cache[0] != sig:
            with torch.no_grad():
                cache = (sig, build())
            self.__dict__["_ewvit_cache"] = cache
        return cache[1]

    def invalidate_native_cache(self):
        self.__dict__.pop("_ewvit_cache", None)

    def _use_native(self, x):
        """Native kernels serve eval-mode CUDA calls.  Training (autograd, BatchNorm batch statistics,
        dropout) and calls with forward hooks installed on sub-modules take the PyTorch composition."""
        if self.training or force_torch_path():
            return False
        if torch.is_grad_enabled() and x.requires_grad:      # gradients w.r.t. the input were asked for
            return False
        if not x.is_cuda:
            from ewvit import EwvitError
            raise EwvitError(f"{type(self).__name__}: eval-mode forward needs CUDA tensors on a B200 "
                             "(the native path has no CPU fallback)")
        for m in self.modules():
            if m._forward_hooks or m._forward_pre_hooks:
                return False
        return True
