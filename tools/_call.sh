mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_s2_9.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_s2_9.log
timeout 300 python bench.py --no-cpu > gpurun_out/bench_s4_1gpu.json 2> gpurun_out/bench_s4_1gpu.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_s4_1gpu.json').read().strip().splitlines()[-1]); print(round(d['value']/1e9,3), round(d['ms_per_step'],3), d['heldout_rmse'], d['e2e']['value']/1e9, d['e2e']['seconds'])"
