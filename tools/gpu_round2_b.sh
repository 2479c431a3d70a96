#!/bin/bash
# round 2, GPU call B: quick parity subset + bench + ncu of the fused kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_pipeline.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -3 gpurun_out/b_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_fused.json 2> gpurun_out/b_bench_fused.err
tail -c 300 gpurun_out/b_bench_fused.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nm_fused -s 60 -c 1 -o gpurun_out/b_fused_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
