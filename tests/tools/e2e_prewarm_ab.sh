# 2-GPU box: the end-to-end figure of bench.py with a short warm-up, several times, with the per-yield gap summary
for run in 1 2 3; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-library-baseline --no-cpu-baseline > gpurun_out/c26_n2_run$run.json 2> gpurun_out/c26_n2_run$run.err; python - <<PY
import json
d=json.loads(open('gpurun_out/c26_n2_run$run.json').read().strip().splitlines()[-1])
print('run $run', 'value', d['value'], 'e2e', round(d['e2e']['value'],2), d['e2e'].get('yield_gaps_ms'))
PY
done
