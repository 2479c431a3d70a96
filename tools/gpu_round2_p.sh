#!/bin/bash
# round 2, GPU call P: final C3 bench line (with cpu baseline) + Stream.run wall clock on the final build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nproc; python -c "import os; print('cpus', os.cpu_count(), len(os.sched_getaffinity(0)))"
timeout 300 python tools/stream_profile.py c3 256 300 2>&1 | grep -E "Stream.run wall|_pipeline.py:.*(upload|run)\)"
timeout 600 python bench.py > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_c3_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'e2e_stream', d['e2e_stream']['value'], d['e2e_stream']['ms_per_step'], 'parity', d.get('parity_checked'), d.get('clocks'))
PY
timeout 300 python bench.py --config default --steps 5 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err; cut -c1-200 gpurun_out/r2_bench_default_n1.json
python __graft_entry__.py smoke 2>&1 | tail -1
