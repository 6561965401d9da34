# ncu evidence on the FINAL code (one GPU): launch list of the bench command + one --set full capture of the dominant kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2_final.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/launches_bench_r2_final.out 2>&1; echo "launch list rc=$?"
# the step's kernels in launch order: probe, scatter<QUAD>, scatter<QUAD_HOT> (returns at once), fold.  -s 8 skips the warm-up steps' scatters
timeout 900 ncu --set full --import-source on --clock-control none -k regex:voxel_scatter_kernel -s 8 -c 2 -o gpurun_out/ncu_voxel_r2_final python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_voxel_r2_final.log 2>&1; echo "ncu voxel rc=$?"
ls -la gpurun_out/*.ncu-rep
