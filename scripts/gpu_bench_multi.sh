# usage: bash scripts/gpu_bench_multi.sh N   — the default bench line on N GPUs of one box (what the driver's SCALE run does)
N=${1:-2}
set -x
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --no-context --no-audit > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err; echo "rc=$?" >> gpurun_out/bench_default_n$N.err
grep -v "Warn\|^  return\|\*\*\*\|OMP" gpurun_out/bench_default_n$N.err | tail -n 6
