"""CPU, world_size 2 over gloo: the host-side logic of the row-band sharding (artstyletransfer_b200/parallel.py)
— band plan, halo sufficiency, packed all-reduce layout, image-gradient sync — with the band arithmetic done by
plain torch on the oracle's VGG (the CUDA kernels are covered by tests/test_gpu_sharding.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from artstyletransfer_b200 import parallel
        from oracle import gatys_oracle as O
        torch.set_num_threads(2)
        parallel.init_sharding()
        assert parallel.world() == (rank, world)
        H, W = 384, 32
        net, cidx, sidx = O.make_vgg19(1234)
        g = torch.Generator().manual_seed(0)
        x = torch.rand((1, 3, H, W), generator=g) * 255 - 120
        tgt = torch.rand((1, 3, H, W), generator=g) * 255 - 120
        plan = parallel.BandPlan(H, rank, world)
        with torch.no_grad():
            full = net(x)
            band = net(x[:, :, plan.lo:plan.hi, :])
            tc_full = net(tgt)[cidx][0]
        chans = [full[k].shape[1] for k in sidx]
        offs, slot, n = parallel.pack_layout(chans)
        packed = torch.zeros(n)
        for i, k in enumerate(sidx):
            a, b, rows = plan.feat_rows(parallel.LAYER_STRIDE[k])
            assert rows == band[k].shape[2]
            fb = band[k][0, :, a:b, :].reshape(chans[i], -1).double()
            packed[offs[i]:offs[i] + chans[i] ** 2] = (fb @ fb.t()).float().reshape(-1)
        ca, cb, _ = plan.feat_rows(parallel.LAYER_STRIDE[cidx])
        ga, gb = plan.global_feat_rows(parallel.LAYER_STRIDE[cidx])
        packed[slot] = ((band[cidx][0, :, ca:cb, :] - tc_full[:, ga:gb, :]).double() ** 2).sum().float()
        parallel._GROUP.all_reduce_sum(packed)
        errs = []
        for i, k in enumerate(sidx):
            f = full[k][0].reshape(chans[i], -1).double()
            ref = (f @ f.t())
            got = packed[offs[i]:offs[i] + chans[i] ** 2].reshape(chans[i], chans[i]).double()
            errs.append(float((got - ref).norm() / ref.norm()))
        sse_ref = float(((full[cidx][0] - tc_full).double() ** 2).sum())
        errs.append(abs(float(packed[slot]) - sse_ref) / sse_ref)
        img = torch.zeros(1, 3, 8, 8, requires_grad=True)
        if rank == 0:
            img.grad = torch.full_like(img, 1.0)       # rank 1 has no gradient yet: must still join
        parallel.sync_image_grad(img)
        out[rank] = (errs, float(img.grad.mean()))
    finally:
        dist.destroy_process_group()


def test_band_sums_equal_full_grams_over_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sorted(out.keys()) == [0, 1]
    for rank in range(world):
        errs, gmean = out[rank]
        assert max(errs) < 1e-5, errs          # halo cropping is exact: only fp32 summation-order noise remains
        assert gmean == 1.0
    assert out[0][0] == out[1][0]              # all ranks hold identical reduced values


def test_init_sharding_requires_a_group():
    from artstyletransfer_b200 import parallel
    parallel.disable_sharding()
    with pytest.raises(RuntimeError, match='process group'):
        parallel.init_sharding()
    assert parallel.world() == (0, 1)
