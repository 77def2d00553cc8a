cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 tests/mgpu_check.py > gpurun_out/mg4.log 2>&1
grep -n "Error\|assert\|Traceback\|mgpu ok" gpurun_out/mg4.log | head -20
grep -n "Traceback" -A25 gpurun_out/mg4.log | head -60
