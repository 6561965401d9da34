mkdir -p gpurun_out
export EVK_TEST_ROUTED=0
( echo 'compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cmax.py tests/test_gpu_image.py tests/test_gpu_voxel.py -q -x -k "onchip or carry or peer_tail or hot or golden or pageable or partial_in_place"'
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cmax.py tests/test_gpu_image.py tests/test_gpu_voxel.py -q -x -k "onchip or carry or peer_tail or hot or golden or pageable or partial_in_place" 2>&1 | tail -4
  echo 'compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_cmax.py tests/test_gpu_image.py -q -x -k "onchip_cells_carry or hot_spot_counts"'
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_cmax.py tests/test_gpu_image.py -q -x -k "onchip_cells_carry or hot_spot_counts" 2>&1 | tail -4
) > gpurun_out/sanitizer_r2.txt 2>&1
cat gpurun_out/sanitizer_r2.txt
