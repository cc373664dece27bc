import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a CPU box: skip them only when deselected
    by marker; if selected without a device they fail loudly inside the product code."""
    return


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture()
def seeded_vgg(monkeypatch):
    """Patch the product's neural_nets.models.vgg19 (hard-coded pretrained=True, like the reference's
    neural_nets.py:19) with a seeded random-init VGG19 — the same shim the goldens were generated with."""
    import torch
    import torchvision
    from artstyletransfer_b200 import neural_nets

    real = torchvision.models.vgg19

    def seeded(pretrained=False, progress=False, **kw):
        torch.manual_seed(1234)
        return real(weights=None)

    monkeypatch.setattr(neural_nets.models, 'vgg19', seeded)
    return seeded


@pytest.fixture(autouse=True)
def cudnn_heuristics_for_parity():
    """Parity tests compare two evaluations of the same convolutions (sharded vs unsharded, graph vs eager, product vs
    oracle): keep cuDNN on its heuristics so both sides run the same engine.  The product default (engine search on,
    feature_path.CUDNN_BENCHMARK) is exercised by test_gpu_closure.py::test_cudnn_engine_search_default."""
    from artstyletransfer_b200 import feature_path
    old = feature_path.CUDNN_BENCHMARK
    feature_path.CUDNN_BENCHMARK = False
    yield
    feature_path.CUDNN_BENCHMARK = old
