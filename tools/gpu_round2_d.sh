#!/bin/bash
# round 2, GPU call D: burst-threshold range split (few rows) -- GPU tests of the burst paths + per-family times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "burst or c4 or rawnorm or raw_normal or stream" > gpurun_out/d_pytest.log 2>&1; tail -3 gpurun_out/d_pytest.log
for a in "c4 32 300" "c4 256 60" "default 256 60" "default 32 300"; do
  timeout 600 python tools/profile_families.py $a >> gpurun_out/d_families.txt 2>&1
done
cat gpurun_out/d_families.txt
