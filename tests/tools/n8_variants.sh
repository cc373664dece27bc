run() { tag=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 --no-library-baseline --no-cpu-baseline $EXTRA > gpurun_out/c17_n8_$tag.json 2> gpurun_out/c17_n8_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c17_n8_$tag.json').read().strip().splitlines()[-1])
    print('$tag', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'loss_first', d['parity']['loss_first'], 'loss_after', d.get('loss_after'))
except Exception as e:
    print('$tag parse failed', e)
PY
}
EXTRA="--profile gpurun_out/c17_prof_n8_d2.txt" run d2 AST_HALO_DEPTH=2
EXTRA="" run d1 AST_HALO_DEPTH=1
run d2_ov03 AST_HALO_DEPTH=2 AST_BAND_OVERHEAD=0.03
run d2_ov005 AST_HALO_DEPTH=2 AST_BAND_OVERHEAD=0.005
run d2_again AST_HALO_DEPTH=2
