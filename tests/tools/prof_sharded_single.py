"""GPU probe: the sharded (halo-exchange, lock-step) schedule on ONE rank under torch.profiler, aten ops grouped by
input shape — finds torch-side launches (copies, fills, strided elementwise) hiding between this library's kernels."""
import os, sys
os.environ['AST_SHARD_SINGLE'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from torch.profiler import ProfilerActivity, profile


class LocalGroup:
    rank, world = 0, 1

    def all_reduce_sum(self, t):
        pass

    def exchange(self, sends, recvs):
        pass


def main():
    args = bench.parse()
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    from artstyletransfer_b200 import neural_style_transfer as nst, parallel
    parallel.init_sharding(LocalGroup())
    bench.seeded_vgg_patch()
    content_levels, style_levels, init, name, _ = bench.build_job(args, dev)
    nst.GRAPH_CLOSURE = False
    job = nst._Job(dev, 'vgg19', style_levels, 'adam', content_levels, init, 10.0, *bench.WEIGHTS, name)
    assert job.pyramid is not None
    for _ in range(3):
        job.optimizer_step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        for _ in range(2):
            job.optimizer_step()
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by='cuda_time_total', row_limit=70,
                                                             max_name_column_width=60, max_shapes_column_width=90))
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=100))


if __name__ == '__main__':
    main()
