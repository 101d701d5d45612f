set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-context --no-audit > gpurun_out/r02x_bench_n2.json 2> gpurun_out/r02x_bench_n2.err; echo "rc=$?" >> gpurun_out/r02x_bench_n2.err
tail -c 800 gpurun_out/r02x_bench_n2.err
