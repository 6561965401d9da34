mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; echo "bench n4 rc=$?"; tail -c 1500 gpurun_out/bench_r2_n4.json; tail -3 gpurun_out/bench_r2_n4.err
