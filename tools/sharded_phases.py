"""Per-phase wall time of the channel-sharded e2e step (run under torchrun with N >= 2 ranks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/sharded_phases.py
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200 import _lib  # noqa: E402
from py_neuromodulation_b200.parallel import ShardedRun, car_shard_factorization, shard_bounds  # noqa: E402
from py_neuromodulation_b200.stream.generator import window_grid  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    n_samples = int(bench.DURATION_S * bench.SFREQ)
    settings = bench.c3_settings()
    c_total = bench.CH_PER_GPU * world
    lo, hi = shard_bounds(c_total, world, rank)
    x = bench.pinned_array(lib, (bench.CH_PER_GPU, n_samples), np.float32)
    bench.synth(bench.CH_PER_GPU, n_samples, seed=rank, out=x)
    channels = get_default_channels_from_data(np.empty((c_total, 1)))
    reref = car_shard_factorization(list(channels["type"]), list(channels["status"]), list(channels["rereference"]), lo, hi)
    dp = nm.DataProcessor(sfreq=bench.SFREQ, settings=settings, channels=channels.iloc[lo:hi].reset_index(drop=True), line_noise=50,
                          verbose=False, device=local, reref_factored=reref)
    starts, lengths, _ = window_grid(n_samples, bench.SFREQ, settings.sampling_rate_features_hz, settings.segment_length_features_ms)
    pipe = dp.plan(int(lengths[0])).pipe
    sh = ShardedRun(pipe, on_gpu=True, shared_host=os.environ.get("SHARED_HOST", "1") == "1")
    n_win = int(starts.size)

    def sync():
        pipe.synchronize()
        torch.cuda.synchronize()

    for it in range(4):
        dist.barrier(); sync()
        t = [time.perf_counter()]
        sh.upload(x); t_host = time.perf_counter() - t[0]; sync() if os.environ.get("SYNC_AFTER_UPLOAD", "0") == "1" else None
        t.append(time.perf_counter())
        sh.run(starts); sync(); t.append(time.perf_counter())
        sh.gather(n_win); sync(); t.append(time.perf_counter())
        dist.barrier(); t.append(time.perf_counter())
        if rank == 0:
            d = np.diff(t) * 1e3
            print(f"iter {it}: upload call returned after {t_host * 1e3:.2f} ms; upload {d[0]:.1f} ms, run {d[1]:.1f} ms, gather {d[2]:.1f} ms, barrier {d[3]:.1f} ms", flush=True)
    dist.barrier()
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
