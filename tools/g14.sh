cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tile_kernel' -s 12 -c 4 -o gpurun_out/r02_tile_v1 -f python bench.py --steps 30 --warmup 5 > gpurun_out/ncu_tile_v1.log 2>&1
tail -3 gpurun_out/ncu_tile_v1.log
ncu -i gpurun_out/r02_tile_v1.ncu-rep --page raw --csv > gpurun_out/tile_v1_raw.csv
python tools/ncu_summary.py gpurun_out/tile_v1_raw.csv
