"""CPU: the C-ABI library loads, exports every symbol include/ast_sm100.h declares, and validates its
arguments (no kernel is launched without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as entry
    entry.build()
    from artstyletransfer_b200 import _lib
    return _lib.load()


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'ast_sm100.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(ast_[a-z0-9_]+)\s*\(', hdr)))


def test_header_symbols_are_exported_and_bound(lib):
    from artstyletransfer_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in ast_sm100.h but not exported'
        assert s in _lib.SIGNATURES, f'{s} has no ctypes signature'
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_sizes(lib):
    from artstyletransfer_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'ast_sm100.h')).read()
    assert lib.ast_version() == _lib.AST_ABI_VERSION == int(re.search(r'#define AST_ABI_VERSION (\d+)', hdr).group(1))
    assert lib.ast_reduce_workspace_bytes() >= 16384
    # workspace covers 148 split-K partials of the largest tile plus the reduce header
    assert lib.ast_gram_workspace_bytes(512, 98304) == 32768 + 148 * 256 * 256 * 4
    assert lib.ast_gram_workspace_bytes(64, 6291456) >= 32768 + 148 * 2 * 64 * 64 * 4
    assert lib.ast_gram_workspace_bytes(0, 10) == 0


def test_argument_validation_without_gpu(lib):
    from artstyletransfer_b200 import _lib
    rc = lib.ast_mse_fwd(None, None, 10, 1.0, None, None, 0, None)
    assert rc == -1 and b'null pointer' in lib.ast_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.ast_bicubic_down2x(p, 3, 7, 8, p, None)
    assert rc == -3 and b'even' in lib.ast_last_error()
    rc = lib.ast_gram_mse_fwd(p, 48, 64, 64, 1.0, None, p, None, p, 1 << 30, 0, None)
    assert rc == -3 and b'multiple of 64' in lib.ast_last_error()
    rc = lib.ast_gram_mse_fwd(p, 64, 64, 64, 1.0, None, p, None, p, 16, 0, None)
    assert rc == -4
    rows = (_lib.HaloRow * 1)()
    assert lib.ast_halo_exchange(rows, 0, None) == -1 and lib.ast_halo_exchange(rows, 17, None) == -1
    assert lib.ast_halo_exchange(rows, 1, None) == -1 and b'halo must be' in lib.ast_last_error()
    rows[0].halo, rows[0].bytes, rows[0].src = p + (-p % 16), 32, p + (-p % 16)
    assert lib.ast_halo_exchange(rows, 1, None) == -1 and b'exchange pointers' in lib.ast_last_error()
    with pytest.raises(RuntimeError, match='ast_mse_bwd failed'):
        _lib.call('ast_mse_bwd', None, None, 1, 1.0, None, None, 0, 0, None)


def test_cpu_tensors_fail_loudly(lib):
    import torch
    from artstyletransfer_b200 import math_utils, ops
    with pytest.raises(RuntimeError, match='CUDA'):
        math_utils.gram_matrix(torch.zeros(1, 64, 4, 4))
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.bicubic_half(torch.zeros(1, 3, 8, 8))


def test_shim_modules_expose_reference_names(lib):
    import importlib
    import sys
    shim = os.path.join(ROOT, 'artstyletransfer_b200', 'shim')
    sys.path.insert(0, shim)
    try:
        for mod in ('config', 'neural_nets', 'math_utils', 'neural_style_transfer'):
            sys.modules.pop(mod, None)
        mu = importlib.import_module('math_utils')
        nn_ = importlib.import_module('neural_nets')
        nst = importlib.import_module('neural_style_transfer')
        cfg = importlib.import_module('config')
    finally:
        sys.path.remove(shim)
        for mod in ('config', 'neural_nets', 'math_utils', 'neural_style_transfer'):
            sys.modules.pop(mod, None)
    for name in ('prepare_model', 'gram_matrix', 'total_variation', 'regularization'):
        assert callable(getattr(mu, name))
    for name in ('Vgg19', 'StyleLoss', 'ContentLoss'):
        assert isinstance(getattr(nn_, name), type)
    for name in ('ContentStylePair', 'RepresentationBuilder', 'LossBuilder', 'NeuralStyleTransfer', 'resize',
                 'neural_style_transfer', 'prepare_img', 'unprepare_img', 'gaussian_mask', 'make_style_noise',
                 'IMAGENET_MEAN_255', 'USE_NORMAL_NOISE_JUST_FOR_DEMONSTRATION',
                 'WITHOUT_GAUSSIAN_MASK_JUST_FOR_DEMONSTRATION', 'SHOW_TEST_IMGS',
                 'IGNORE_GRADIENT_MAP_JUST_FOR_DEMONSTRATION'):
        assert hasattr(nst, name), name
    c = cfg.Config()
    assert (c.content_weight, c.style_weight, c.tv_weight, c.optimizer, c.levels_num, c.iters_num) == \
           (1e3, 4e5, 1e2, 'lbfgs', 2, 500)
    assert cfg.simultaneous_tasks_count == 2
    # every field of the reference's Config (config.py:5-30), in its positional order, with its default
    assert list(vars(c).items()) == [
        ('content_weight', 1e3), ('style_weight', 4e5), ('tv_weight', 1e2), ('optimizer', 'lbfgs'), ('model', 'vgg19'),
        ('init_method', 'content+noise'), ('levels_num', 2), ('iters_num', 500), ('noise_factor', 0.95),
        ('noise_levels', (9, 18, 36, -1, 0)), ('noise_levels_central_amplitude', (0.30, 0.20, 0.10, 0.20, 0.20)),
        ('noise_levels_peripheral_amplitude', (0.20, 0.30, 0.40, 0.10, 0.00)),
        ('noise_levels_dispersion', (0.20, 0.30, 0.40, 0.60, 0.30))]
    c2 = cfg.Config(2.0, optimizer='adam', noise_levels=(-1,))
    assert (c2.content_weight, c2.style_weight, c2.optimizer, c2.noise_levels) == (2.0, 4e5, 'adam', (-1,))
    with pytest.raises(TypeError):
        cfg.Config(no_such_setting=1)
    with pytest.raises(ValueError):
        mu.prepare_model('alexnet', 'cpu')


def test_process_rejects_unknown_optimizer_and_cpu(lib):
    import asyncio
    import numpy as np
    from artstyletransfer_b200 import neural_style_transfer as nst

    async def run(device, opt):
        drv = nst.NeuralStyleTransfer(device, 'vgg19', [np.zeros((32, 32, 3), np.float32)], opt)
        async for _ in drv.process([np.zeros((32, 32, 3), np.float32)], np.zeros((32, 32, 3), np.float32), 10.0, 1,
                                   1e3, 4e5, 1e2, 'x'):
            pass

    with pytest.raises(RuntimeError, match='CUDA only'):
        asyncio.run(run('cpu', 'adam'))
