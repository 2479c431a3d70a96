#!/bin/bash
# Final measurements of a round on ONE B200 (run through gpurun): tests, bench lines, ncu launch list, ncu --set full of the C3 kernels.
# Outputs go to gpurun_out/<tag>_*; large ncu reports stay in /tmp on the box.
tag=${1:-r2}
out=gpurun_out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -2 $out/${tag}_pytest_gpu.log
python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_c3_n1.json 2> $out/${tag}_bench_c3_n1.err; cut -c1-160 $out/${tag}_bench_c3_n1.json
python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; cut -c1-200 $out/${tag}_bench_reference_arm.json
python bench.py --config default --steps 5 > $out/${tag}_bench_default_n1.json 2> $out/${tag}_bench_default_n1.err; cut -c1-160 $out/${tag}_bench_default_n1.json
python bench.py --config c4 --steps 3 > $out/${tag}_bench_c4_n1.json 2> $out/${tag}_bench_c4_n1.err; cut -c1-160 $out/${tag}_bench_c4_n1.json
python bench.py --config c5 --steps 2 > $out/${tag}_bench_c5_n1.json 2> $out/${tag}_bench_c5_n1.err; cut -c1-160 $out/${tag}_bench_c5_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_ncu_launches.csv python bench.py --quick --no-cpu-baseline --no-parity --steps 2 --warmup 3 > $out/${tag}_ncu_launches_bench.json 2> /dev/null
ncu --set full --import-source on --clock-control none -k regex:"nm_convx_kernel|nm_notchx_kernel|nm_specx_kernel|nm_prep_kernel" -s 30 -c 4 -o /tmp/${tag}_c3 python tools/profile_families.py c3 256 20 > $out/${tag}_ncu_c3.log 2>&1
ncu -i /tmp/${tag}_c3.ncu-rep --page raw --csv > $out/${tag}_ncu_c3_raw.csv
python tools/ncu_summary.py $out/${tag}_ncu_c3_raw.csv > $out/${tag}_ncu_c3_summary.txt
ncu -i /tmp/${tag}_c3.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${tag}_c3_lines.csv
python tools/ncu_lines.py /tmp/${tag}_c3_lines.csv 25 > $out/${tag}_ncu_c3_lines.txt
ls -la $out/${tag}_*
