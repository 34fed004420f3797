mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_s2_6.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_s2_6.log
grep -n "^E  *AssertionError\|^E  *assert" gpurun_out/pytest_s2_6.log | head
B="python bench.py --no-e2e --no-cpu --steps 10 --warmup 3"
run() { tag=$1; shift; env "$@" timeout 120 $B > gpurun_out/s2f_bench_$tag.json 2> gpurun_out/s2f_bench_$tag.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/s2f_bench_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value']/1e9,3), round(d['ms_per_step'],3), d['heldout_rmse'])
except Exception as e: print('$tag ERR', open('gpurun_out/s2f_bench_$tag.err').read()[-300:])"; }
run l8d4 X=1
run l8d2 MFSGD_HDEPTH=2
run l16d4 MFSGD_HOT_LANES=16
run l16d2 MFSGD_HOT_LANES=16 MFSGD_HDEPTH=2
run l32d4 MFSGD_HOT_LANES=32
run l8d4c1 MFSGD_HOT_CTAS=1
run l16d4c2 MFSGD_HOT_LANES=16 MFSGD_HOT_CTAS=2
run l8d4nopdl MFSGD_PDL=0
timeout 120 python tools/run_workload.py ml20m > gpurun_out/s2f_ml20m.log 2>&1; tail -c 330 gpurun_out/s2f_ml20m.log
timeout 300 python tools/run_workload.py powerlaw 8 > gpurun_out/s2f_powerlaw8.log 2>&1; tail -c 700 gpurun_out/s2f_powerlaw8.log
