cd $GRAFT_REPO_ROOT
export HB2_BENCH_CPU=0
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r02_b_2gpu_v13.json 2> gpurun_out/r02_b_2gpu_v13.err
python - <<P
import json; d=json.load(open('gpurun_out/r02_b_2gpu_v13.json')); print('2gpu', d['value'], round(d['ms_per_step'],4), d.get('phases'), d.get('parity',{}).get('pass'), d.get('strong_baseline'), d.get('e2e'))
P
