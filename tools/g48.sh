cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 tests/mgpu_check.py > gpurun_out/mg2.log 2>&1
grep -n "Error\|mgpu ok" gpurun_out/mg2.log | head -8
grep -n "Traceback" -A22 gpurun_out/mg2.log | head -30
export HB2_BENCH_REPL=5,5,5 HB2_BENCH_CPU=0 HB2_BENCH_STRONG=0
for p in 1 0 1 0; do
HB2_P2P=$p timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2955$p bench.py --gpus 2 --steps 300 --warmup 30 > gpurun_out/r02_mg2_p2p$p.json 2> gpurun_out/r02_mg2_p2p$p.err
python - <<P
import json
f='gpurun_out/r02_mg2_p2p$p.json'
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print('p2p=$p', d['value'], d['ms_per_step'], d['phases']['integrate_ms_per_step'], d['phases']['force_ms_per_step'], d.get('parity',{}).get('pass'))
except Exception as e: print(f, 'ERR', e)
P
done
