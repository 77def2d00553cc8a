cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 tests/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
HB2_BENCH_CPU=0 HB2_BENCH_STRONG=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 100 --warmup 20 > gpurun_out/r02_bench_4gpu_64M_final.json 2> gpurun_out/r02_bench_4gpu_64M_final.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_4gpu_64M_final.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'], d['parity']['pass'], d['clocks'])
P
