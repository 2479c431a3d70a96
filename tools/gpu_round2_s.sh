#!/bin/bash
# round 2, GPU call S: compute-sanitizer (memcheck, then racecheck) over the tests of the code added in the last session
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests -m gpu -x -q \
   -k "featnorm or rawnorm_minmax or rawnorm_robust or sklearn_feature or across_chunks or (three_times and f64) or burst_thresholds_incremental" \
   > gpurun_out/s_racecheck.log 2>&1
echo "exit $?"; tail -6 gpurun_out/s_racecheck.log
