mkdir -p gpurun_out
R="timeout 100 python tools/run_workload.py ml100k 1"
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "rounds",d["rounds"],"hot",d["hot_items"],"rmse",[round(x,4) for x in d["heldout_rmse_per_epoch"][:6]],"...",round(d["heldout_rmse_per_epoch"][-1],4),"launches",d["launches"][0], "ms", round(d["epoch_ms"][-1],3))
except Exception as e:
    print(sys.argv[1],"ERR",e, open(sys.argv[1]).read()[-400:])
PY
}
RW_TAG=_a $R > gpurun_out/e_a.log 2>&1; show gpurun_out/e_a.log
MFSGD_PDL=2 RW_TAG=_b $R > gpurun_out/e_b.log 2>&1; show gpurun_out/e_b.log
MFSGD_PDL=0 RW_TAG=_c $R > gpurun_out/e_c.log 2>&1; show gpurun_out/e_c.log
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_s2_5.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_s2_5.log
grep -n "^E  *AssertionError" gpurun_out/pytest_s2_5.log | head
MFSGD_PDL=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "rmse_parity" > gpurun_out/pytest_s2_5b.log 2>&1; echo "pytest pdl2 rc=$?"
tail -8 gpurun_out/pytest_s2_5b.log
grep -n "^E  *AssertionError" gpurun_out/pytest_s2_5b.log | head
for P in 0 1; do MFSGD_PDL=$P timeout 120 python bench.py --no-e2e --no-cpu --steps 10 --warmup 3 > gpurun_out/s2e_bench_pdl$P.json 2> gpurun_out/s2e_bench.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/s2e_bench_pdl$P.json').read().strip().splitlines()[-1]); print('pdl$P', round(d['value']/1e9,3), round(d['ms_per_step'],3), d['heldout_rmse'])"; done
