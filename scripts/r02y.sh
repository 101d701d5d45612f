set -x
timeout 600 python -m pytest tests/test_loss_gpu.py -q -x > gpurun_out/r02y_loss.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_loss.log
tail -n 3 gpurun_out/r02y_loss.log
timeout 300 python scripts/ial_diag.py > gpurun_out/r02y_ial_diag.log 2>&1
grep golden gpurun_out/r02y_ial_diag.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload c5_train > gpurun_out/r02y_c5_n2.json 2> gpurun_out/r02y_c5_n2.err; echo "rc=$?" >> gpurun_out/r02y_c5_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload c1_train > gpurun_out/r02y_c1_n2.json 2>> gpurun_out/r02y_c5_n2.err; echo "rc=$?" >> gpurun_out/r02y_c5_n2.err
grep -v Warn gpurun_out/r02y_c5_n2.err | tail -n 5
