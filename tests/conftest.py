import os
import sys

import pytest

os.environ.setdefault("EWVIT_ALLOW_RANDOM_BACKBONE", "1")     # tests use key-addressed weights, never ImageNet ones

TESTS = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(TESTS)
PKG = os.path.join(REPO, "efficient-wavelet-vit_b200")
for p in (PKG, REPO, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # an eval-mode call that silently drops to the PyTorch composition (grad mode left on) must fail a test, not pass slowly
    config.addinivalue_line("filterwarnings", "error:.*builds a graph through the PyTorch composition")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(TESTS, "golden", "ewvit_golden.pt"), weights_only=False)


@pytest.fixture(scope="session")
def manifest():
    from _weights import load_manifest
    return load_manifest()


@pytest.fixture(scope="session")
def dama_sd(manifest):
    """Key-addressed fp32 weights for the dynamic-mode path (dama.* + classifier.*)."""
    from _weights import state_dict_from_manifest
    return state_dict_from_manifest(manifest, seed=0, prefixes=("dama.", "classifier."))
