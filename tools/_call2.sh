#!/bin/bash
# 2-GPU box: the ring window transport against ncclSend/ncclRecv (parity + throughput), and a Yahoo-shaped-per-visit ring of 2
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e --no-traffic"
run() { tag=$1; shift; ( env "$@" timeout 200 $TR --master-port $((29500 + RANDOM % 400)) $B > gpurun_out/r2g_bench2_$tag.json 2> gpurun_out/r2g_bench2_$tag.err ); echo "$tag rc=$?" >> gpurun_out/r2g_status.log; }
rm -f gpurun_out/r2g_status.log
run window MFSGD_TRACE=1
run nccl MFSGD_RING_TRANSPORT=nccl
run window_waitkernel MFSGD_RING_WAIT=kernel
run window_sigwrite MFSGD_RING_SIGNAL=write
run window_lanes1 MFSGD_LANES=1
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --tb=short -k "ring" 2>&1 | tail -25 ) > gpurun_out/r2g_pytest_multi2.log 2>&1
# Yahoo-shaped visits on a ring of 2: 450 K users x 34 K items, 44 M ratings, 2 sub-stripes x 2 rounds x 2 sub-shards
Y="RW_SHAPE=450000,34000,44000000 RW_STRIPES=2 RW_ROUNDS=2 RW_SHARDS=2"
yrun() { tag=$1; shift; ( env $Y "$@" timeout 200 $TR --master-port $((29500 + RANDOM % 400)) tests/mp/ring_workload.py yahoo 4 > gpurun_out/r2g_yahoo2_$tag.log 2>&1 ); echo "yahoo2 $tag rc=$?" >> gpurun_out/r2g_status.log; }
yrun window
yrun nccl MFSGD_RING_TRANSPORT=nccl
yrun window_nopdl MFSGD_PDL=0
yrun window_lanes1 MFSGD_LANES=1
echo done
