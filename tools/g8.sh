cd $GRAFT_REPO_ROOT
export HB2_BENCH_REPL=5,5,5 HB2_BENCH_DEBUG=1 HB2_BENCH_CPU=0 HB2_BENCH_STRONG=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02_mg2b.json 2> gpurun_out/r02_mg2b.err
grep -v "^\[rebuild\]" gpurun_out/r02_mg2b.err | tail -12
cat gpurun_out/r02_mg2b.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['parity'], d['strong_baseline'], d['clocks'])"
unset HB2_BENCH_REPL HB2_BENCH_DEBUG HB2_BENCH_STRONG
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -15
