mkdir -p gpurun_out/r2i
NG=${NG:-4}
timeout 300 python -m pytest tests/test_dist_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 --skip strong_256 > gpurun_out/r2i/bench_n$NG.json 2> gpurun_out/r2i/bench_n$NG.err; echo "rc=$?"; tail -3 gpurun_out/r2i/bench_n$NG.err
