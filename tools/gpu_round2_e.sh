#!/bin/bash
# round 2, GPU call E: 1536 / 3072-point 'same'-mode banks (12 x R1 x 16 plans) -- GPU tests, A/B against the power-of-two plans,
# 3 vs 4 resident CTAs per SM for the 1536-point bank kernel, occupancy target of the burst-threshold range split
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/e
timeout 1200 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; tail -3 ${o}_pytest.log
run() { echo "== $1" >> ${o}_families.txt; shift; timeout 600 "$@" >> ${o}_families.txt 2>&1; }
run "c3 mixed" python tools/profile_families.py c3 256 300
NMB200_MIXED_RADIX=0 run "c3 pow2" python tools/profile_families.py c3 256 300
run "default mixed" python tools/profile_families.py default 256 60
NMB200_MIXED_RADIX=0 run "default pow2" python tools/profile_families.py default 256 60
run "c5-like mixed (128 ch @ 2 kHz)" python tools/profile_families.py default 128 60 2000
NMB200_MIXED_RADIX=0 run "c5-like pow2" python tools/profile_families.py default 128 60 2000
for occ in 1 2 3 4; do NMB200_BURST_SPLIT_OCC=$occ run "c4 share split occ $occ" python tools/profile_families.py c4 32 300; done
timeout 600 python bench.py --quick --no-cpu-baseline > ${o}_bench_c3.json 2> ${o}_bench_c3.err; cut -c1-300 ${o}_bench_c3.json
cp py_neuromodulation_b200/csrc/libnmb200.so /tmp/lib_default.so
cp gpurun_tmp/libnmb200_bk4.so py_neuromodulation_b200/csrc/libnmb200.so
run "c3 mixed, 4 CTAs/SM bank" python tools/profile_families.py c3 256 300
run "default mixed, 4 CTAs/SM bank" python tools/profile_families.py default 256 60
cp /tmp/lib_default.so py_neuromodulation_b200/csrc/libnmb200.so
cat ${o}_families.txt
