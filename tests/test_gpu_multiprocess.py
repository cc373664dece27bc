"""GPU, >= 2 devices: the row-band sharded closure over REAL NCCL ranks (one process per GPU, torch.distributed.run)
against the unsharded closure — the multi-process counterpart of tests/test_gpu_sharding.py's one-GPU emulation.
Skipped on a one-GPU box (NCCL refuses two ranks on one device); `gpurun --gpus 2` runs it."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(nproc, env_extra=None, args=(), port=29533, timeout=420):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={nproc}',
           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tests', 'tools', 'nccl_parity.py'),
           *args]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    return r, (json.loads(lines[-1]) if lines else None)


def need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f'needs {n} GPUs (have {torch.cuda.device_count()})')


@pytest.mark.timeout(600)
@pytest.mark.parametrize('nproc', [2, 4])
def test_nccl_ranks_match_unsharded_closure(nproc):
    need(nproc)
    r, res = launch(nproc, port=29533 + nproc)
    assert r.returncode == 0 and res is not None and res['ok'], (r.stdout[-3000:], r.stderr[-3000:])
    assert res['world'] == nproc and res['loss_rel_err'] <= 1e-4 and res['grad_rel_err'] <= 2e-3


@pytest.mark.timeout(600)
def test_peer_memory_halo_matches_unsharded_closure():
    """AST_HALO=peer: halo rows pushed through NVLink peer memory by ast_halo_exchange instead of NCCL send/recv."""
    need(2)
    r, res = launch(2, {'AST_HALO': 'peer'}, port=29541)
    assert r.returncode == 0 and res is not None and res['ok'], (r.stdout[-3000:], r.stderr[-3000:])
    assert res['halo'] == 'peer'
