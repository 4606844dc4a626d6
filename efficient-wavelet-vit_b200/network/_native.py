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


class NativeMixin:
    """Caches a native runner per module.  The runner holds re-laid-out COPIES of the weights (folded BatchNorm, bf16,
    tap-major), so it is rebuilt whenever a parameter/buffer on the path was modified in place, replaced or moved:
    the cache key lists (key, data_ptr, version counter, device) per tensor, and every ``train()``/``eval()`` transition
    and every ``load_state_dict`` drops it.  One write path is invisible to all of that: ``p.data.copy_()`` /
    ``p.data.fill_()`` bump neither the version nor the pointer (EMA/SWA swaps, ``vector_to_parameters``); after such a
    write in eval mode call ``invalidate_native_cache()`` (or toggle ``train()``/``eval()``)."""

    def _native_tensors(self):
        return self.state_dict(keep_vars=True).items()

    def _native_signature(self):
        return tuple((k, t.data_ptr(), t._version, str(t.device)) for k, t in self._native_tensors())

    def _native_runner(self, build):
        sig = self._native_signature()
        cache = self.__dict__.get("_ewvit_cache")
        if cache is None or cache[0] != sig:
            with torch.no_grad():
                cache = (sig, build())
            self.__dict__["_ewvit_cache"] = cache
        return cache[1]

    def invalidate_native_cache(self):
        """Drop the native runner of this module and of every native sub-module."""
        for m in self.modules():
            m.__dict__.pop("_ewvit_cache", None)

    def train(self, mode=True):
        self.__dict__.pop("_ewvit_cache", None)          # nn.Module.train recurses, so sub-modules drop theirs too
        return super().train(mode)

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_ewvit_cache", None)
        return super()._load_from_state_dict(*args, **kwargs)

    def _use_native(self, x):
        """Native kernels serve eval-mode CUDA calls that need no autograd graph.  Training (BatchNorm batch statistics,
        dropout), calls that need gradients (grad mode on and the input or any parameter on the path requires grad: the
        native outputs carry no ``grad_fn``) and calls with forward hooks installed on sub-modules take the PyTorch
        composition of the same modules."""
        if self.training:
            return False
        if not x.is_cuda:
            from ewvit import EwvitError
            raise EwvitError(f"{type(self).__name__}: eval-mode forward needs CUDA tensors on a B200 "
                             "(the native path has no CPU fallback)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            _warn_grad_mode_once(type(self).__name__)
            return False
        for m in self.modules():
            if m._forward_hooks or m._forward_pre_hooks:
                return False
        return True


_WARNED = set()


def _warn_grad_mode_once(name):
    if name not in _WARNED:
        _WARNED.add(name)
        import warnings
        warnings.warn(f"{name}: eval-mode forward with autograd enabled and trainable parameters builds a graph through the "
                      "PyTorch composition; wrap inference in torch.no_grad() (as eval.py:149 does) to run the native "
                      "sm_100a kernels", stacklevel=3)
