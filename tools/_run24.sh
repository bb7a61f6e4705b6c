OUT=gpurun_out/r2F; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; head -c 300 $OUT/bench.json; echo; tail -3 $OUT/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; head -c 300 $OUT/bench_ref.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --skip A,C,D,strong_256 --no-cpu --steps 2 --warmup 3 --pairs 30 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
