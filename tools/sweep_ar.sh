#!/bin/bash
for c in 8 16 32 0; do
  if [ "$c" = "0" ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$c; fi
  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 tools/time_allreduce.py 2>&1 | grep "all-reduce" 
done
