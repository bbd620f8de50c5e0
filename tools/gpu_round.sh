#!/bin/bash
# full GPU suite, then UAV slab-vs-scalar timing, then the headline bench
mkdir -p gpurun_out
bash tools/gpu_tests.sh 2>&1 | tail -n 12
rm -f gpurun_out/uav_slab.jsonl
bash tools/gpu_uav_slab.sh 2>&1 | grep -v "^\.\|passed" | tail -n 12
python bench.py --steps 300 --warmup 20 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('K1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
