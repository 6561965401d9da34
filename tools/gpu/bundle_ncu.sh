# ncu evidence for the round-2 kernels (one GPU)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/launches_bench_r2.out 2>&1; echo "launch list rc=$?"
N=50000000 REPS=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:voxel_routed -s 1 -c 1 -o gpurun_out/ncu_routed_r2f python tools/prof_routed.py > gpurun_out/ncu_routed_r2f.log 2>&1; echo "ncu routed rc=$?"
SIZES=50000000 timeout 900 ncu --set full --clock-control none -k regex:cmax_onchip -s 3 -c 1 -o gpurun_out/ncu_onchip_f_r2 python tools/bench_cmax.py > gpurun_out/ncu_onchip_f_r2.log 2>&1; echo "ncu onchip f-only rc=$?"
timeout 900 python tools/profile_hot.py > gpurun_out/profile_hot_r2.log 2>&1; echo "hot rc=$?"; tail -5 gpurun_out/profile_hot_r2.log
