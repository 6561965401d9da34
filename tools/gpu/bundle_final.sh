mkdir -p gpurun_out
timeout 60 python tools/try_routed.py > gpurun_out/try_routed_final.log 2>&1; rc=$?; echo "try_routed rc=$rc"; tail -4 gpurun_out/try_routed_final.log
if [ $rc -eq 0 ]; then export EVK_TEST_ROUTED=1; else export EVK_TEST_ROUTED=0; fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$? (EVK_TEST_ROUTED=$EVK_TEST_ROUTED)"; tail -6 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2_n1_f.json 2> gpurun_out/bench_r2_n1_f.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r2_n1_f.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_r2_ref.json
