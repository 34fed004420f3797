mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s2_8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_s2_8.log
timeout 300 python bench.py > gpurun_out/bench_s3_1gpu.json 2> gpurun_out/bench_s3_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/bench_s3_reference.json 2> gpurun_out/bench_s3_reference.err; echo "ref rc=$?"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
python -c "
import json
d=json.loads(open('gpurun_out/bench_s3_1gpu.json').read().strip().splitlines()[-1]); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), d['heldout_rmse'], d['e2e']['value']/1e9, d['e2e']['seconds'], d['roofline'].get('l2_bound'), d['clocks'])
d=json.loads(open('gpurun_out/bench_s3_reference.json').read().strip().splitlines()[-1]); print(d['value']/1e6, d['cpu_baseline']['cores'])"
