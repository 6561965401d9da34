mkdir -p gpurun_out
timeout 120 python tools/try_routed.py > gpurun_out/try_routed_r2e.log 2>&1; echo "try_routed rc=$?"; tail -6 gpurun_out/try_routed_r2e.log
timeout 1300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r2a.log
ONLY=skip timeout 600 python tools/bench_variants.py > gpurun_out/variants_r2a.log 2>&1; echo "variants rc=$?"; grep -E "zipf|uniform" gpurun_out/variants_r2a.log | head -40
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1_a.json 2> gpurun_out/bench_r2_n1_a.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_r2_n1_a.json
