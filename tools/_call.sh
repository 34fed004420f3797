mkdir -p gpurun_out; rm -f gpurun_out/s3_proxy.jsonl
for c in 0 48 64 96 128 192 256; do PROXY_SHARDS=16 timeout 60 python tools/ring8_proxy.py 0 $c 8 >> gpurun_out/s3_proxy.jsonl 2>&1; done
PROXY_SHARDS=16 MFSGD_HDEPTH=2 timeout 60 python tools/ring8_proxy.py 0 96 8 >> gpurun_out/s3_proxy.jsonl 2>&1
PROXY_SHARDS=16 MFSGD_HOT_CTAS=3 timeout 60 python tools/ring8_proxy.py 0 96 8 >> gpurun_out/s3_proxy.jsonl 2>&1
PROXY_SHARDS=16 MFSGD_HOT_LANES=16 timeout 60 python tools/ring8_proxy.py 0 96 8 >> gpurun_out/s3_proxy.jsonl 2>&1
PROXY_SHARDS=8 timeout 60 python tools/ring8_proxy.py 0 96 8 >> gpurun_out/s3_proxy.jsonl 2>&1
PROXY_SHARDS=8 timeout 60 python tools/ring8_proxy.py 0 128 8 >> gpurun_out/s3_proxy.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/s3_proxy.jsonl'):
    try:
        d=json.loads(l); print(d['shards'], d['hot_chunk'], round(d['epoch_ms'],3), round(d['gupdates_s'],2), d['launches'])
    except Exception: print(l[:200])
PY
