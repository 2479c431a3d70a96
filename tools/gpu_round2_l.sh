#!/bin/bash
# round 2, GPU call L: L2 prefetch of the next item in front of the shared-memory epilogues (sharp waves, burst envelopes); 3 vs 4 CTAs per SM
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/l
timeout 600 python -m pytest tests -m gpu -x -q -k "burst or sharp or c4 or default" > ${o}_pytest.log 2>&1; tail -2 ${o}_pytest.log
run() { echo "== $1" >> ${o}_families.txt; shift; timeout 600 "$@" >> ${o}_families.txt 2>&1; }
run "default, prefetch" python tools/profile_families.py default 256 60
run "c5-like, prefetch" python tools/profile_families.py default 128 60 2000
cp py_neuromodulation_b200/csrc/libnmb200.so /tmp/lib_default.so
cp gpurun_tmp/libnmb200_b1_3.so py_neuromodulation_b200/csrc/libnmb200.so
run "default, no prefetch, single-filter kernels at 3 CTAs/SM (168 registers)" python tools/profile_families.py default 256 60
cp /tmp/lib_default.so py_neuromodulation_b200/csrc/libnmb200.so
grep -E "^==|device time|sharpwave|burst_envelope|notch" ${o}_families.txt
