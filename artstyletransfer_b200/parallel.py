"""Row-band sharding of a pyramid level across the GPUs of one box.  Filled in below; with a
single process (no torch.distributed group) every hook is a no-op and the path is exactly the 1-GPU one."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def maybe_shard(loss_builders, optimizing_img, neural_net, content_idx, style_idx, weights):
    return None


def sync_image_grad(optimizing_img):
    return None
