#!/bin/bash
# round 2, GPU call A: parity of the fused kernel + bench fused / un-fused + ncu
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -3 gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_fused.json 2> gpurun_out/a_bench_fused.err
tail -c 600 gpurun_out/a_bench_fused.json
NMB200_FUSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_unfused.json 2> gpurun_out/a_bench_unfused.err
tail -c 300 gpurun_out/a_bench_unfused.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_fused -s 60 -c 1 -o gpurun_out/a_fused_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_full.log 2>&1
ls -la gpurun_out | tail
