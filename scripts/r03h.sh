set -x
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-context --no-audit > gpurun_out/r03h_bench_n2.json 2> gpurun_out/r03h_bench_n2.err; echo "rc=$?" >> gpurun_out/r03h_bench_n2.err
grep -v "Warn\|^  return\|\*\*\*\|OMP" gpurun_out/r03h_bench_n2.err | tail -n 6
python scripts/profile_eval.py > gpurun_out/r03h_eval_profile_1m.txt 2>/dev/null
head -n 40 gpurun_out/r03h_eval_profile_1m.txt
