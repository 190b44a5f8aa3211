#!/bin/bash
N=$1
run() { echo "== $*"; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/exp_allreduce.py 2>/dev/null | tail -1; }
run X=default
run NCCL_ALGO=NVLS
run NCCL_ALGO=Ring
run NCCL_ALGO=Tree
run NCCL_MIN_NCHANNELS=32
run NCCL_ALGO=Ring NCCL_MIN_NCHANNELS=32
run NCCL_NVLS_ENABLE=0
