#!/bin/bash
# 8-GPU box: scaling bench (8, 4, 2, 1), configs[3]/[4] on their own configuration, multi-GPU parity tests.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_env.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 240 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu > gpurun_out/r2e_bench_${n}gpu.json 2> gpurun_out/r2e_bench_${n}gpu.err
done
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu --no-traffic > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/r2e_bench_1gpu.err
for w in yahoo powerlaw; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --workload $w --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2e_bench_${w}_8gpu.json 2> gpurun_out/r2e_bench_${w}_8gpu.err
done
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short 2>&1 | tail -40 ) > gpurun_out/r2e_pytest_multi.log 2>&1
echo done
