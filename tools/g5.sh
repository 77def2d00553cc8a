cd $GRAFT_REPO_ROOT
HB2_BENCH_CPU=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_force_kernel|nlist_build_kernel' -s 30 -c 4 -o gpurun_out/r02_rows_v0 python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_rows_v0.log 2>&1
tail -2 gpurun_out/ncu_rows_v0.log
