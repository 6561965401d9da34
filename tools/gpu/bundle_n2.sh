# two GPUs: the multi-GPU tests (skipped on one GPU) and the 2-GPU bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_voxel.py::test_peer_reduced_voxel_two_gpus tests/test_gpu_cmax.py::test_peer_cmax_two_gpus -x -q > gpurun_out/pytest_n2_r2.log 2>&1; echo "pytest n2 rc=$?"; tail -15 gpurun_out/pytest_n2_r2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"; tail -c 3000 gpurun_out/bench_r2_n2.json; tail -5 gpurun_out/bench_r2_n2.err
