#!/bin/bash
for c in 1 0; do echo "--- STB_VOLUME_CACHEL=$c"; STB_VOLUME_CACHEL=$c timeout 900 python bench.py --steps 8 --warmup 3 --no-extras --no-train --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], {k:round(v['ms_total']/8,3) for k,v in d['kernels'].items()})"; done
STB_VOLUME_CACHEL=0 timeout 600 python -m pytest tests/test_split_gpu.py -m gpu -q -k "gwcnet" 2>&1 | tail -2
