cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -25
export HB2_BENCH_REPL=5,5,5 HB2_BENCH_CPU=0 HB2_BENCH_STRONG=0
for leg in 1 0; do
echo "---- HB2_COMM_LEGACY=$leg"
HB2_COMM_LEGACY=$leg HB2_TRACE_REBUILD=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$leg bench.py --gpus 2 --steps 40 --warmup 10 > /dev/null 2> gpurun_out/r02_mg2_trace_leg$leg.err
grep "^\[rebuild\]" gpurun_out/r02_mg2_trace_leg$leg.err | tail -8 | head -5
HB2_COMM_LEGACY=$leg timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$leg bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_mg2_8M_leg$leg.json 2> gpurun_out/r02_mg2_8M_leg$leg.err
python - <<P
import json
f='gpurun_out/r02_mg2_8M_leg$leg.json'
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'], d.get('parity',{}).get('pass'))
except Exception as e: print(f, 'ERR', e)
P
done
