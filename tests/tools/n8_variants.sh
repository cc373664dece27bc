# 8-GPU box: the bench line at N = 8 (twice), one NCCL protocol variant, and N = 4.  Usage: gpurun --gpus 8 -- bash tests/tools/n8_variants.sh
run() { n=$1; tag=$2; shift 2; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 30 --warmup 5 --no-library-baseline --no-cpu-baseline > gpurun_out/c22_n${n}_$tag.json 2> gpurun_out/c22_n${n}_$tag.err; echo "$tag rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c22_n${n}_$tag.json').read().strip().splitlines()[-1])
    print('N=$n $tag', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', round(d['e2e']['value'],1), 'loss_first', d['parity']['loss_first'], 'loss_after', d.get('loss_after'))
except Exception as e:
    print('$tag parse failed', e)
PY
}
run 8 default AST_HALO_DEPTH=2
run 8 ll128 NCCL_PROTO=LL128
run 4 default AST_HALO_DEPTH=2
run 8 again AST_HALO_DEPTH=2
