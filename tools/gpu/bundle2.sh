mkdir -p gpurun_out
timeout 100 python tools/try_routed.py > gpurun_out/try_routed_r2f.log 2>&1; rc=$?; echo "try_routed rc=$rc"; tail -14 gpurun_out/try_routed_r2f.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_r2b.log
if [ $rc -eq 0 ]; then
  timeout 900 python tools/bench_variants.py > gpurun_out/variants_r2b.log 2>&1; echo "variants rc=$?"; grep -E "voxel|zipf" gpurun_out/variants_r2b.log | head -60
fi
