mkdir -p gpurun_out
for N in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_s2_${N}gpu.json 2> gpurun_out/bench_s2_${N}gpu.err; echo "bench$N rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_s2_${N}gpu.json').read().strip().splitlines()[-1]); print($N, round(d['value']/1e9,3), round(d['ms_per_step'],3), d['heldout_rmse'], round(d['e2e']['value']/1e9,2), d['e2e']['seconds'], d['config']['stripes_per_gpu'], d['config']['shards_per_gpu'], d['config']['rounds'], d['breakdown_ms_per_step_rank0'])"
done
