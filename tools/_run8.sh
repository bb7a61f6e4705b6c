mkdir -p gpurun_out/r2j
NG=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/r2j/bench_n$NG.json 2> gpurun_out/r2j/bench_n$NG.err; echo "rc=$?"; tail -3 gpurun_out/r2j/bench_n$NG.err
nvidia-smi topo -m > gpurun_out/r2j/topo.txt 2>&1
