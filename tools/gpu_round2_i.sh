#!/bin/bash
# round 2, GPU call J: sorted-queue warp-per-row burst thresholds
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/j
timeout 900 python -m pytest tests -m gpu -x -q -k "burst or c4 or rawnorm or raw_normal or stream" > ${o}_pytest.log 2>&1; tail -2 ${o}_pytest.log
run() { echo "== $1" >> ${o}_families.txt; shift; timeout 600 "$@" >> ${o}_families.txt 2>&1; }
run "default" python tools/profile_families.py default 256 60
run "c4 share, split" python tools/profile_families.py c4 32 300
NMB200_BURST_SPLIT_OCC=0 run "c4 share, no split" python tools/profile_families.py c4 32 300
NMB200_BURST_SPLIT_OCC=1 run "c4 share, split occ 1" python tools/profile_families.py c4 32 300
run "c5-like" python tools/profile_families.py default 128 60 2000
grep -E "^==|device time|burst_threshold|burst thresholds" ${o}_families.txt
