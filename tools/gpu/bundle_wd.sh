#!/bin/bash
# watchdog check: the routed gate, then the routed tests incl. the fault injection
timeout 300 python tools/try_routed.py > gpurun_out/try_routed_wd.log 2>&1; echo "try_routed rc=$?"; tail -5 gpurun_out/try_routed_wd.log
timeout 600 python -m pytest tests/test_gpu_voxel.py -m gpu -q -k "routed or watchdog or probe or vs_oracle or golden" > gpurun_out/pytest_wd.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_wd.log
