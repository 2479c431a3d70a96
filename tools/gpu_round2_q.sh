#!/bin/bash
# round 2, GPU call Q: final sanity of the last build (GPU tests, smoke) + Stream.run / bench of the default feature set with the per-chunk normaliser
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest.log 2>&1; tail -3 gpurun_out/q_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/stream_profile.py default 256 300 2>&1 | grep -E "Stream.run wall|_pipeline.py:.*(upload|run)\)"
timeout 300 python bench.py --config default --steps 5 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().splitlines()[-1])
print('default value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'e2e_stream', d.get('e2e_stream', {}).get('ms_per_step'), 'parity', d.get('parity_checked'))
PY
