#!/bin/bash
# round 2, 8 GPUs: strong-scaling configurations C4 / C5 with the final build (the driver runs the C3 scaling series itself)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in c4 c5; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --config $cfg --steps 3 --warmup 3 \
      > gpurun_out/r2_bench_${cfg}_n8.json 2> gpurun_out/r2_bench_${cfg}_n8.err
  tail -c 600 gpurun_out/r2_bench_${cfg}_n8.json | cut -c1-600; echo
done
