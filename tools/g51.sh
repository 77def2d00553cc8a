cd $GRAFT_REPO_ROOT
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 tests/p2p8_check.py > gpurun_out/p2p8.log 2>&1
grep -n "p2p8 ok\|Error" gpurun_out/p2p8.log | head -5
grep -n "Traceback" -A15 gpurun_out/p2p8.log | head -20
