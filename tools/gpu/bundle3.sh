mkdir -p gpurun_out
timeout 60 python tools/try_routed.py > gpurun_out/try_routed_r2g.log 2>&1; rc=$?; echo "try_routed rc=$rc"; tail -14 gpurun_out/try_routed_r2g.log
if [ $rc -eq 0 ]; then export EVK_TEST_ROUTED=1; else export EVK_TEST_ROUTED=0; fi
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$? (EVK_TEST_ROUTED=$EVK_TEST_ROUTED)"; tail -12 gpurun_out/pytest_gpu_r2c.log
timeout 300 python tools/profile_hot.py > gpurun_out/profile_hot_r2.log 2>&1; echo "hot rc=$?"; cat gpurun_out/profile_hot_r2.log
