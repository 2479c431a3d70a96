#!/bin/bash
# round 2, GPU call R: final sanity of the last build -- GPU tests, smoke, default bench line (C3)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r_pytest.log 2>&1; tail -3 gpurun_out/r_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r_bench_c3.json 2> gpurun_out/r_bench_c3.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r_bench_c3.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_stream', d['e2e_stream']['ms_per_step'], 'parity', d.get('parity_checked'), 'launches', d.get('gpu_launches'))
PY
