#!/bin/bash
# like ab_run.sh, with the cost-kernel parity tests
mkdir -p gpurun_out
for n in "$@"; do
  export FSGM_LIB=/root/repo/scratch/ab_$n.so
  timeout 300 python -m pytest tests/test_epi_gpu.py -m gpu -x -q -k "stages_and_gateway or wild_geometry or full_kitti or config_a or wave_pipeline" > gpurun_out/ab_${n}_pytest.log 2>&1
  echo "$n pytest: $(tail -1 gpurun_out/ab_${n}_pytest.log)"
  python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ab_${n}_bench.json 2> gpurun_out/ab_${n}_bench.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${n}_bench.json"))
print("$n", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
PY
  python bench.py --steps 10 --warmup 3 --no-cpu --no-overlap > gpurun_out/ab_${n}_bench_no.json 2> /dev/null
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${n}_bench_no.json"))
print("$n no-overlap", round(d["value"],1), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
PY
done
