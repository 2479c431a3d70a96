"""Where does the channel-sharded e2e step spend its time?  (torchrun, N >= 2 ranks, collectives inside libnmb200)

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sharded_phases.py [steps]

C3 weak scaling (256 channels per rank).  Every variant is timed like bench.py's e2e (barrier + device sync on both sides, wall clock,
max over ranks, mean of `steps` steps):
  resident       group sums + all-reduce + re-reference + window kernels, recording already in HBM (= bench `value`)
  upload         sliced H2D of the shard + per-slice group sums + per-slice ncclAllReduce + re-reference, no window kernels
  upload+run     the above + all window kernels, results stay on the device
  e2e            the above + rows of every finished chunk copied into the shared page-locked matrix (= bench `e2e`)
  e2e solo       e2e with the ranks taking turns (one rank at a time: no contention for host memory / PCIe / NVLink)
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200 import _lib  # noqa: E402
from py_neuromodulation_b200.parallel import NativeComm, ShardedRun, car_shard_factorization, shard_bounds  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    lib = _lib.load()
    comm = NativeComm.from_env(device=local)
    n_samples = int(bench.DURATION_S * bench.SFREQ)
    settings = bench.c3_settings()
    c_total = bench.CH_PER_GPU * world
    lo, hi = shard_bounds(c_total, world, rank)
    x = bench.pinned_array(lib, (hi - lo, n_samples), np.float32)
    bench.synth_rows(lo, hi - lo, n_samples, seed=0, out=x)
    channels = get_default_channels_from_data(np.empty((c_total, 1)))
    reref = car_shard_factorization(list(channels["type"]), list(channels["status"]), list(channels["rereference"]), lo, hi)
    dp = nm.DataProcessor(sfreq=bench.SFREQ, settings=settings, channels=channels.iloc[lo:hi].reset_index(drop=True), line_noise=50,
                          verbose=False, device=local, reref_factored=reref)
    starts, lengths, _ = window_grid(n_samples, bench.SFREQ, settings.sampling_rate_features_hz, settings.segment_length_features_ms)
    pipe = dp.plan(int(lengths[0])).pipe
    sh = ShardedRun(pipe, on_gpu=True, comm=comm)
    n_win = int(starts.size)

    def upload_only():
        sh.upload(x)
        _lib.check(lib.nm_prepare_resident_sharded(pipe._h, comm._h))  # waits for every slice's reduction, re-references all of it

    def upload_run():
        sh.upload(x)
        pipe.run(starts, download=False)

    def e2e():
        sh.upload(x)
        sh.run(starts)
        sh.gather(n_win)

    def resident():
        _lib.check(lib.nm_prepare_resident_sharded(pipe._h, comm._h))
        pipe.run(starts, download=False)

    def timed(fn, n):
        pipe.synchronize(); comm.barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        pipe.synchronize()
        return comm.max((time.perf_counter() - t0) * 1e3 / n)

    for _ in range(3):
        e2e()
    res = {}
    for name, fn in (("resident", resident), ("upload", upload_only), ("upload+run", upload_run), ("e2e", e2e)):
        res[name] = timed(fn, steps)
    # ranks take turns: the other ranks still take part in the collectives (a rank cannot all-reduce alone), so "solo" is
    # measured without the exchange -- plain upload + run + download of ONE rank's shard while the others idle
    solo = 0.0
    plain = nm.DataProcessor(sfreq=bench.SFREQ, settings=settings, channels=channels.iloc[lo:hi].reset_index(drop=True), line_noise=50,
                             verbose=False, device=local).plan(int(lengths[0])).pipe
    out = bench.pinned_array(lib, (n_win, plain.F), np.float64)
    plain.upload(x); plain.run(starts, out=out)
    for r in range(world):
        comm.barrier()
        if r == rank:
            t0 = time.perf_counter()
            for _ in range(steps):
                plain.upload(x)
                plain.run(starts, out=out)
            solo = (time.perf_counter() - t0) * 1e3 / steps
    comm.barrier()
    res["e2e solo (no exchange, one rank at a time)"] = comm.max(solo)
    if rank == 0:
        print(f"C3 weak scaling, {world} ranks x 256 channels, {n_win} windows, mean of {steps} steps, max over ranks, ms per step:")
        for k, v in res.items():
            print(f"  {k:45s} {v:8.2f}")
    comm.barrier()
    sh.close()
    comm.close()


if __name__ == "__main__":
    main()
