#!/bin/bash
# One gpurun call of the usual per-change sequence on a single B200: GPU parity tests, smoke, bench lines; optional extras by flag.
#   tools/gpu_round.sh TAG [tests] [bench] [ref] [sanitize] [ncu]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for what in "$@"; do
  case $what in
    tests)    timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log ;;
    smoke)    timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log ;;
    bench)    timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 3000 $OUT/bench.json; tail -3 $OUT/bench.err ;;
    benchB)   timeout 600 python bench.py --skip A,C,D,strong_256 --no-cpu > $OUT/benchB.json 2> $OUT/benchB.err; echo "benchB rc=$?"; head -c 2500 $OUT/benchB.json; tail -3 $OUT/benchB.err ;;
    ref)      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; head -c 600 $OUT/bench_ref.json ;;
    sanitize) for tool in memcheck racecheck; do
                timeout 1500 compute-sanitizer --tool $tool --log-file $OUT/sanitizer_$tool.log python tools/sanitize_cases.py > $OUT/sanitizer_$tool.out 2>&1
                echo "$tool rc=$?"; tail -3 $OUT/sanitizer_$tool.log; tail -2 $OUT/sanitizer_$tool.out
              done ;;
    ncu)      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --skip A,C,D,strong_256 --no-cpu --steps 2 --warmup 3 --pairs 30 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
