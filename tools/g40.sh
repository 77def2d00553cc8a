cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -25
HB2_BENCH_CPU=0 timeout 300 python bench.py --steps 300 --warmup 50 > gpurun_out/r02_b_v18.json 2> gpurun_out/r02_b_v18.err; python - <<P
import json; d=json.load(open('gpurun_out/r02_b_v18.json')); print('v18', round(d['ms_per_step'],4), d['clocks'])
P
