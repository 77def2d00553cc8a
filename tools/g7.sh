cd $GRAFT_REPO_ROOT
export HB2_BENCH_REPL=5,5,5 HB2_BENCH_DEBUG=1 HB2_BENCH_CPU=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_mg2.json 2> gpurun_out/r02_mg2.err
grep -v "^\[rebuild\]" gpurun_out/r02_mg2.err | tail -15
cat gpurun_out/r02_mg2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'])"
HB2_TRACE_REBUILD=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 40 --warmup 10 > /dev/null 2> gpurun_out/r02_mg2_trace.err
grep "^\[rebuild\]" gpurun_out/r02_mg2_trace.err | tail -24
