#!/bin/bash
# round 2, GPU call O: deferred (threaded, double-buffered) staging of pageable recordings -- GPU tests + nm.Stream.run wall clock A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/o
timeout 1200 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; tail -3 ${o}_pytest.log
for d in 1 0; do
  echo "== NMB200_DEFERRED_UPLOAD=$d" >> ${o}_stream.txt
  NMB200_DEFERRED_UPLOAD=$d timeout 300 python tools/stream_profile.py c3 256 300 2>&1 | grep -E "Stream.run wall|_pipeline.py:.*(upload|run)\)" >> ${o}_stream.txt
  NMB200_DEFERRED_UPLOAD=$d timeout 300 python tools/stream_profile.py default 256 300 2>&1 | grep -E "Stream.run wall|_pipeline.py:.*(upload|run)\)" >> ${o}_stream.txt
done
cat ${o}_stream.txt
timeout 600 python bench.py --no-cpu-baseline > ${o}_bench_c3.json 2> ${o}_bench_c3.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/o_bench_c3.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'e2e_stream', d.get('e2e_stream'), 'parity', d.get('parity_checked'))
PY
