#!/bin/bash
# round 2, GPU call C: bank kernel with the late filter-spectrum prefetch (un-fused), default feature set fused vs un-fused
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NMB200_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_unfused.json 2> gpurun_out/c_bench_unfused.err
tail -c 300 gpurun_out/c_bench_unfused.json
for f in 0 1; do
  echo "== NMB200_FUSED=$f" >> gpurun_out/c_families.txt
  NMB200_FUSED=$f timeout 600 python tools/profile_families.py default 256 60 >> gpurun_out/c_families.txt 2>&1
  NMB200_FUSED=$f timeout 600 python tools/profile_families.py c3 256 60 >> gpurun_out/c_families.txt 2>&1
  NMB200_FUSED=$f timeout 600 python tools/profile_families.py c4 32 300 >> gpurun_out/c_families.txt 2>&1
done
cat gpurun_out/c_families.txt
