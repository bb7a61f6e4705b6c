mkdir -p gpurun_out/r2o
timeout 1200 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pyd -c 120 --csv --log-file gpurun_out/r2o/conflicts_pyd.csv python bench.py --skip A,D,strong_256 --no-cpu --steps 1 --warmup 1 --pairs 15 > gpurun_out/r2o/ncu_conf.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/r2o/ncu_conf.log | cut -c1-300
