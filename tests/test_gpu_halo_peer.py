"""GPU: ast_halo_exchange in loop-back on ONE device — two emulated neighbours A and B whose entries travel in the
same launch, each one's "remote" pointers aimed at the other's staging slots and counters.  Exercises the whole protocol (push, release/acquire counters, slot parity,
self-resetting tickets, count advance) over several consecutive exchanges (passed on a B200 in round 1); the
two-process run over NVLink is bench.py with AST_HALO=peer under torchrun (pending: round 2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(60)
@pytest.mark.parametrize('row_floats', [4, 1000, 3072 * 64])
def test_halo_exchange_loopback(row_floats):
    from artstyletransfer_b200 import _lib as L, ops
    dev = torch.device('cuda', 0)
    nbytes = row_floats * 4
    slot = (nbytes + 255) // 256 * 256
    stage = torch.zeros(2, 2 * slot // 4, device=dev)                 # [A, B] x slot pair
    flags = torch.zeros(2, 64, dtype=torch.int32, device=dev)
    state = torch.zeros(2, 64, dtype=torch.int32, device=dev)
    band = [torch.zeros(4, row_floats, device=dev) for _ in range(2)]   # rows: halo, edge, edge, halo
    border = torch.full((row_floats,), 7.0, device=dev)
    for it in range(5):
        for k in range(2):
            band[k][1:3] = torch.randn(2, row_floats, device=dev) + 10 * it + k
        rows = (L.HaloRow * 3)()
        # A's bottom edge <-> B's top edge
        for k, (src, halo) in enumerate(((band[0][2], band[0][3]), (band[1][1], band[1][0]))):
            o = 1 - k
            rows[k].src, rows[k].halo, rows[k].bytes, rows[k].slot_stride = src.data_ptr(), halo.data_ptr(), nbytes, slot
            rows[k].dst_remote, rows[k].flag_remote = stage[o].data_ptr(), flags[o].data_ptr()
            rows[k].stage, rows[k].flag_local, rows[k].state = stage[k].data_ptr(), flags[k].data_ptr(), state[k].data_ptr()
        rows[2].halo, rows[2].bytes = border.data_ptr(), nbytes        # src NULL: zero a border halo
        ops._launch(dev, ('halo_exchange', 3), 'ast_halo_exchange', rows, 3)
        torch.cuda.synchronize()
        assert torch.equal(band[0][3], band[1][1]) and torch.equal(band[1][0], band[0][2])
        assert float(border.abs().max()) == 0.0
        assert state[:, 0].tolist() == [it + 1, it + 1] and flags[:, 0].tolist() == [it + 1, it + 1]
        assert state[:, 1:3].abs().max().item() == 0
        border.fill_(7.0)
