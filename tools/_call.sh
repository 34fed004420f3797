mkdir -p gpurun_out
rm -f gpurun_out/l2_peak_all.jsonl
for mb in 61 30 100; do timeout 60 tools/l2_peak $mb >> gpurun_out/l2_peak_all.jsonl; done
head -1 gpurun_out/l2_peak_all.jsonl > gpurun_out/l2_peak.json; cat gpurun_out/l2_peak_all.jsonl
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_workloads.py -m gpu -q 2>&1 | tail -3
