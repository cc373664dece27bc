"""CPU: the parts of bench.py's contract that need no GPU — both arms state the same workload, the metric name follows
the level count, the work accounting knows every kernel key the product emits."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def args(**kw):
    base = dict(gpus=1, steps=20, warmup=5, impl='ours', levels=4, optimizer='adam', precision=None, init='structured')
    base.update(kw)
    return argparse.Namespace(**base)


def test_both_arms_state_the_same_workload():
    a = bench.workload_config(args())
    b = bench.workload_config(args(impl='reference'))
    assert a == b and a['levels'] == 4 and a['image'] == [2048, 3072]
    assert 'BASELINE configs[3]' in a['workload'] and a['optimizer'] == 'adam'
    assert bench.workload_config(args(levels=2, init='pixel'))['init'].startswith('content+noise, PIXEL_WIDE')
    assert 'configs[0]' in bench.workload_config(args(levels=1, optimizer='lbfgs'))['workload']


def test_metric_name_and_init_args():
    assert bench.metric_name(4) == 'pyramid_L3_optim_steps_per_s'
    assert bench.metric_name(3) == 'pyramid_L2_optim_steps_per_s'
    assert bench.init_args(args())[2] == (9, 18, 36, -1, 0)            # config.Config() defaults
    assert bench.init_args(args(init='pixel'))[1:3] == (0.5, (-1,))      # lab.py PIXEL_WIDE_NOISE_CONFIG


def test_kernel_work_accounting():
    c, hw = 64, 6291456
    assert bench.kernel_work(('gram_fwd_nhwc', c, hw)) == (4.0 * c * hw + 8.0 * c * c, 2.0 * c * c * hw)
    assert bench.kernel_work(('gram_bwd_nhwc', c, hw, 3))[0] == 12.0 * c * hw + 4.0 * c * c
    assert bench.kernel_work(('gram_bwd_nhwc_bf16', 512, 98304, 1))[1] == 2.0 * 512 * 512 * 98304
    assert bench.kernel_work(('down2x_tv', 3, 2048, 3072))[0] == 15.0 * 2048 * 3072
    assert bench.kernel_work(('noise_init', 2048, 3072))[0] == 24.0 * 2048 * 3072
    assert bench.kernel_work(('unprepare', 2048 * 3072))[0] == 24.0 * 2048 * 3072
    assert bench.kernel_work(('allreduce_packed_grams', 10)) == (0.0, 0.0)
    assert bench.kernel_work(('tv_bwd_rows', 3 * 352 * 3072)) == bench.kernel_work(('tv_bwd', 3 * 352 * 3072))
    assert bench.kernel_work(('tv_bwd_rows', 12))[0] > 0
    assert bench.pixel_ratio(4, 2) == 17.0


def test_yield_gap_summary():
    """e2e diagnostics: gaps between consecutive yields of the timed region (the first `warmup` yields are skipped)."""
    times = [0.0, 1.0, 2.0, 2.010, 2.020, 2.080, 2.090]              # warm-up 3: timed gaps 10, 10, 60, 10 ms
    g = bench._gap_summary(times, 3)
    assert g == {'median': 10.0, 'max': 60.0, 'max_at_timed_step': 3, 'over_3x_median': 1}
    assert bench._gap_summary([0.0, 1.0], 3) is None
