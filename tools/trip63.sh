#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/t63_bench_n8.json 2> gpurun_out/t63_bench_n8.err; echo "bench N=8 rc=$?"; tail -2 gpurun_out/t63_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t63_bench_n8.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches'):
    print(k, d.get(k))
print('train_step', {k:d['train_step'].get(k) for k in ('value','n_gpus','ms_per_step','allreduce_ms','allreduce_alone_ms','allreduce_busbw_gbs','grad_disagreement_after_allreduce','error')})
PY
