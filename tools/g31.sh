cd $GRAFT_REPO_ROOT
export HB2_BENCH_REPL=5,5,5 HB2_BENCH_CPU=0 HB2_BENCH_STRONG=0
HB2_TRACE_REBUILD=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 10 > gpurun_out/r02_mg2_trace.json 2> gpurun_out/r02_mg2_trace.err
grep "^\[rebuild\]" gpurun_out/r02_mg2_trace.err | tail -18
echo ---- 1 GPU same system
HB2_TRACE_REBUILD=1 timeout 300 python bench.py --gpus 1 --steps 40 --warmup 10 > gpurun_out/r02_sg_trace.json 2> gpurun_out/r02_sg_trace.err
grep "^\[rebuild\]" gpurun_out/r02_sg_trace.err | tail -8
echo ---- untraced 2 GPU
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_mg2_8M.json 2> gpurun_out/r02_mg2_8M.err
python - <<'P'
import json
for f in ('gpurun_out/r02_mg2_8M.json','gpurun_out/r02_sg_trace.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'])
    except Exception as e: print(f, 'ERR', e)
P
tail -5 gpurun_out/r02_mg2_8M.err
