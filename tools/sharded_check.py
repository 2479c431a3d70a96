"""Value check of the channel-sharded path on real GPUs (run under torchrun with N >= 2 ranks):
the merged result of the sharded run must equal the un-sharded run of the same recording on rank 0's GPU.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.parallel import ShardedRun, car_shard_factorization, merge_permutation, shard_bounds  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    native = "--native" in sys.argv  # collectives inside libnmb200 (nm_comm_*) instead of torch.distributed
    comm = None
    torch.cuda.set_device(local)
    if native:
        from py_neuromodulation_b200.parallel import NativeComm

        comm = NativeComm.from_env(device=local)
    else:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    c_total, T = 10 * world + 3, 90_000  # uneven shards, >= 65 536 samples: sliced upload with per-slice all-reduce
    rng = np.random.default_rng(7)
    x = (np.cumsum(rng.standard_normal((c_total, T)), axis=1) * 0.01 + rng.standard_normal((c_total, T))).astype(np.float32)
    ok = True
    for name, settings in (("c3", None), ("default", nm.NMSettings.get_default())):
        if settings is None:
            settings = nm.NMSettings.get_default().reset()
            for f in ("fft", "bandpass_filter", "raw_hjorth", "linelength"):
                settings.features[f] = True
        settings.postprocessing.feature_normalization = False
        lo, hi = shard_bounds(c_total, world, rank)
        channels = get_default_channels_from_data(x)
        reref = car_shard_factorization(list(channels["type"]), list(channels["status"]), list(channels["rereference"]), lo, hi)
        dp = nm.DataProcessor(sfreq=1000, settings=settings, channels=channels.iloc[lo:hi].reset_index(drop=True), line_noise=50,
                              verbose=False, device=local, reref_factored=reref)
        starts, lengths, _ = window_grid(T, 1000, settings.sampling_rate_features_hz, settings.segment_length_features_ms)
        for shared in (True, False):
            sh = ShardedRun(dp.plan(1000).pipe, on_gpu=True, shared_host=shared, comm=comm)
            for rep in range(2):
                dp.plan(1000).pipe.reset_state()
                sh.upload(x[lo:hi])
                sh.run(starts)
                got = sh.gather(len(starts))
            if rank == 0:
                cols, perm = merge_permutation(settings, list(channels["new_name"]), 1000, 1000, world)
                merged = np.array(got[:, perm])
                full = nm.DataProcessor(sfreq=1000, settings=settings, channels=channels, line_noise=50, verbose=False, device=local)
                ref_cols, ref = full.process_windows(x, starts, 1000)
                assert list(ref_cols) == list(cols)
                err = np.nanmax(np.abs(merged - ref) / np.maximum(np.abs(ref), 1.0))
                same_nan = np.array_equal(np.isnan(merged), np.isnan(ref))
                print(f"{name} shared_host={shared}: {len(starts)} windows x {len(cols)} columns, max rel err vs un-sharded run {err:.2e}, "
                      f"NaN pattern equal: {same_nan}", flush=True)
                ok = ok and err < 1e-9 and same_nan
            comm.barrier() if native else dist.barrier()
            sh.close()
    if rank == 0:
        print("SHARDED CHECK", "PASSED" if ok else "FAILED", flush=True)
    if native:
        comm.close()
    else:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
