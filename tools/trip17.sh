#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t17_bench_n2.json 2> gpurun_out/t17_bench_n2.err; echo "bench N=2 rc=$?"; tail -3 gpurun_out/t17_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t17_bench_n2.json').read().strip().splitlines()[-1])
for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches'):
    print(k, d.get(k))
print('train_step', {k:d['train_step'].get(k) for k in ('value','n_gpus','ms_per_step','allreduce_ms','allreduce_busbw_gbs','allreduce_share_of_step','grad_disagreement_after_allreduce','error')})
print('keys', sorted(d.keys()))
PY
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "follow_the_tensor_device" -rs > gpurun_out/t17_dev.log 2>&1; tail -3 gpurun_out/t17_dev.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 | cut -c1-400
