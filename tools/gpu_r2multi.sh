#!/bin/bash
# Run on an N-GPU box (gpurun --gpus N): GPU tests incl. NCCL + peer ones, bench at 1 GPU, both arms' N-GPU bench
# with the peer-memory gather (default) and the NCCL all_gather.
# Usage: tools/gpu_r2multi.sh <tag> <ngpus> [steps]
TAG=${1:-r2}
N=${2:-2}
STEPS=${3:-5}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_$TAG.txt 2>&1
timeout 900 python -W ignore -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps $STEPS --warmup 3 > $OUT/bench_${TAG}_g1.json 2> $OUT/bench_${TAG}_g1.err
tail -c 2500 $OUT/bench_${TAG}_g1.json; tail -3 $OUT/bench_${TAG}_g1.err
for PEER in 1 0; do
PVK_PEER_GATHER=$PEER timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$PEER \
    bench.py --gpus $N --steps $STEPS --warmup 3 --no-e2e > $OUT/bench_${TAG}_g${N}_peer$PEER.json 2> $OUT/bench_${TAG}_g${N}_peer$PEER.err
echo "rc=$?"; tail -c 900 $OUT/bench_${TAG}_g${N}_peer$PEER.json; grep -v "^\[" $OUT/bench_${TAG}_g${N}_peer$PEER.err | tail -4
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2> $OUT/bench_${TAG}_reference.err
tail -c 600 $OUT/bench_${TAG}_reference.json
