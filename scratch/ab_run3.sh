#!/bin/bash
mkdir -p gpurun_out
for n in "$@"; do
  export FSGM_LIB=/root/repo/scratch/ab_$n.so
  python bench.py --steps 10 --warmup 3 --no-cpu --no-overlap 2>/dev/null > gpurun_out/ab_${n}_no.json
  python -c "
import json
d=json.load(open('gpurun_out/ab_${n}_no.json'))
print('$n no-overlap', round(d['value'],1), {k:round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
timeout 200 python -m pytest tests/test_epi_gpu.py -m gpu -x -q -k "stages_and_gateway or full_kitti or cluster_fast" 2>&1 | tail -1
