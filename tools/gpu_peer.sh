#!/bin/bash
# Run on a 2+-GPU box (gpurun --gpus N): the NCCL tests incl. the peer-memory gather, then the bench at N GPUs
# with the NCCL all_gather and with the fused rename + peer-store gather.
# Usage: tools/gpu_peer.sh <tag> <ngpus> [steps]
TAG=${1:-r2}
N=${2:-2}
STEPS=${3:-5}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
PVK_TEST_PEER=1 timeout 600 python -W ignore -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -15 | tee $OUT/pytest_multi_$TAG.log
for PEER in 0 1; do
PVK_PEER_GATHER=$PEER timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$PEER \
    bench.py --gpus $N --steps $STEPS --warmup 3 --no-e2e > $OUT/bench_${TAG}_g${N}_peer$PEER.json 2> $OUT/bench_${TAG}_g${N}_peer$PEER.err
tail -c 700 $OUT/bench_${TAG}_g${N}_peer$PEER.json; grep -v "^\[" $OUT/bench_${TAG}_g${N}_peer$PEER.err | tail -4
done
