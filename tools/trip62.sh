#!/bin/bash
timeout 900 python -m pytest tests/test_split_gpu.py tests/test_gpu_umma.py tests/test_update_umma_gpu.py tests/test_igev_stereo_gpu.py -m gpu -q -x 2>&1 | tail -3
for c in 1 0; do echo "--- STB_UMMA_CSPLIT16=$c"; STB_UMMA_CSPLIT16=$c timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-train 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['epe_e2e_px'], {k:round(v['ms_total']/10,3) for k,v in d['kernels'].items()}); [print(n, v) for n,v in d['layers'].items() if n.startswith('2d') and ('128->128' in n or '32->32' in n)]"
STB_UMMA_CSPLIT16=$c timeout 400 python tools/model_bench.py --model raft --height 512 --width 1024 --iters 32 --cuda-graph 2>&1 | tail -1 | sed 's/.*"ms_per_forward": \([0-9.]*\).*/raft \1 ms/'
done
