#!/bin/bash
# round 2, GPU call M: full GPU test suite with the scikit-learn feature normalisers; cost of the O(n^2) methods at 256 channels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
o=gpurun_out/m
timeout 1200 python -m pytest tests -m gpu -x -q > ${o}_pytest.log 2>&1; tail -3 ${o}_pytest.log
python - > ${o}_norm_cost.txt 2>&1 <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import py_neuromodulation_b200 as nm
from py_neuromodulation_b200.stream.generator import window_grid
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data
x = np.random.default_rng(0).random((256, 60000), dtype=np.float32)
for method in ("zscore", "median", "minmax", "robust", "quantile"):
    s = nm.NMSettings.get_default()
    s.feature_normalization_settings.normalization_method = method
    dp = nm.DataProcessor(sfreq=1000, settings=s, channels=get_default_channels_from_data(x), line_noise=50, verbose=False)
    starts, lengths, _ = window_grid(x.shape[1], 1000, s.sampling_rate_features_hz, s.segment_length_features_ms)
    pipe = dp.plan(1000).pipe
    pipe.upload(x)
    pipe.run(starts, download=False); pipe.synchronize(); pipe.reset_state()
    pipe.set_profiling(True)
    pipe.run(starts, download=False); pipe.synchronize()
    prof = pipe.profile()
    print(method, "normalizer %.3f ms per %d windows x %d columns" % (prof["normalizer"][0], starts.size, pipe.F))
PY
cat ${o}_norm_cost.txt
