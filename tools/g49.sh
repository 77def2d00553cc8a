cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29504 tests/mgpu_check.py > gpurun_out/mg4.log 2>&1
grep -n "Error\|mgpu ok" gpurun_out/mg4.log | head -8
grep -n "Traceback" -A22 gpurun_out/mg4.log | head -30
