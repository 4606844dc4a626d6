"""Stand-in for the third-party ``efficientnet_pytorch`` package.  TEST INFRASTRUCTURE ONLY.

Lets the UNMODIFIED reference ``network/sfe.py`` (``from efficientnet_pytorch import
EfficientNet``, line 4) import in the build container, where the real package is absent.
Only ``tests/golden/make_golden.py`` puts this directory on ``sys.path``.  The b0 itself is
the repo's own restatement with upstream-compatible key names; the dynamic-mode path that
the golden vectors cover never executes it (it only has to construct).
"""
import importlib.util
import os

_here = os.path.dirname(os.path.abspath(__file__))
_src = os.path.normpath(os.path.join(_here, "..", "..", "..", "efficient-wavelet-vit_b200", "network", "_effnet_b0.py"))
_spec = importlib.util.spec_from_file_location("_ewvit_effnet_b0", _src)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
EfficientNet = _mod.EfficientNet
