#!/bin/bash
# Run on an 8-GPU box (gpurun --gpus 8): NCCL tests and the bench at N = 4 and 8 (both arms at 8).
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -W ignore -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee $OUT/pytest_multi_$TAG.log
for N in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${TAG}_g$N.json 2> $OUT/bench_${TAG}_g$N.err
tail -c 600 $OUT/bench_${TAG}_g$N.json; tail -2 $OUT/bench_${TAG}_g$N.err
done
