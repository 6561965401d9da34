mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "bench n8 rc=$?"; tail -c 2500 gpurun_out/bench_r2_n8.json; tail -3 gpurun_out/bench_r2_n8.err
