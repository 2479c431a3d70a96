"""Channel-sharded execution with world_size 2 over gloo on the CPU: the host logic
(shard bounds, global common-average factorisation, all-reduce of the group sums, gather, column merge) runs
against the thread-emulated kernels and must reproduce the oracle on the un-sharded recording."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import neural_like


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, x: np.ndarray, settings_dict: dict, out_file: str, shared_host: bool) -> None:
    import torch.distributed as dist

    from tests.emu_support import load_emu
    from py_neuromodulation_b200 import _lib

    _lib._LIB = load_emu()
    import py_neuromodulation_b200 as nm
    from py_neuromodulation_b200.parallel import ShardedRun, car_shard_factorization, merge_permutation, shard_bounds
    from py_neuromodulation_b200.stream.generator import window_grid
    from py_neuromodulation_b200.utils.channels import get_default_channels_from_data

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        settings = nm.NMSettings(**settings_dict)
        c_total = x.shape[0]
        lo, hi = shard_bounds(c_total, world, rank)
        channels = get_default_channels_from_data(x)
        reref = car_shard_factorization(list(channels["type"]), list(channels["status"]), list(channels["rereference"]), lo, hi)
        dp = nm.DataProcessor(sfreq=1000, settings=settings, channels=channels.iloc[lo:hi].reset_index(drop=True), line_noise=50,
                              verbose=False, reref_factored=reref)
        starts, lengths, _ = window_grid(x.shape[1], 1000, settings.sampling_rate_features_hz, settings.segment_length_features_ms)
        plan = dp.plan(int(lengths[0]))
        run = ShardedRun(plan.pipe, on_gpu=False, shared_host=shared_host)
        for _ in range(2):  # second pass: buffers / shared matrix reused
            run.upload(x[lo:hi].astype(np.float32))
            run.run(starts)
            gathered = run.gather(len(starts))
        if rank == 0:
            cols, perm = merge_permutation(settings, list(channels["new_name"]), 1000, int(lengths[0]), world)
            np.savez(out_file, cols=np.array(cols), mat=gathered[:, perm])
        dist.barrier()
        run.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shared_host,n_samples,rate", [(True, 1800, 10), (False, 1800, 10), (True, 66_000, 2)],
                         ids=["shared-host-matrix", "nccl-style-gather", "sliced-upload"])
def test_two_rank_channel_shard_matches_oracle(tmp_path, shared_host, n_samples, rate):
    import torch.multiprocessing as mp

    import py_neuromodulation_b200 as nm
    from oracle import np_oracle as orc
    from tests.emu_support import build_emu

    build_emu()
    x = neural_like(31, 5, n_samples)  # odd channel count: shards of 3 and 2; >= 65 536 samples: upload in 8 reduced slices
    s = nm.NMSettings.get_default().reset()
    s.sampling_rate_features_hz = rate
    s.features.fft = True
    s.features.raw_hjorth = True
    s.features.bandpass_filter = True
    s.features.linelength = True
    out_file = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(2, _free_port(), x, s.model_dump(), out_file, shared_host), nprocs=2, join=True)
    got = np.load(out_file)
    ref_cols, ref = orc.run_offline(x, 1000, s.model_dump())
    ref_cols, ref = ref_cols[:-1], ref[:, :-1]  # drop the time column
    assert list(got["cols"]) == ref_cols
    err = np.abs(got["mat"] - ref) / np.maximum(np.abs(ref), 1.0)
    assert err.max() < 1e-9, err.max()


def test_shard_bounds_and_factorisation():
    from py_neuromodulation_b200.parallel import car_shard_factorization, shard_bounds

    assert [shard_bounds(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_bounds(256, 8, r) for r in range(8)][-1] == (224, 256)
    types = ["ecog"] * 4 + ["seeg"] * 3
    status = ["good", "good", "bad", "good", "good", "good", "good"]
    refs = ["average", "average", "average", "None", "average", "average", "average"]
    g, group_of, gcoef, ptr, col, diag = car_shard_factorization(types, status, refs, 0, 7)
    assert g == 2
    assert list(group_of) == [0, 0, -1, 0, 1, 1, 1]
    # ecog: good channels 0, 1, 3 -> channel 0 is referenced to the mean of {1, 3}
    assert gcoef[0, 0] == -0.5 and diag[0] == 1.5 and gcoef[2].tolist() == [0, 0] and diag[2] == 1 and gcoef[3].tolist() == [0, 0]
    assert gcoef[4, 1] == -0.5 and diag[4] == 1.5
    with pytest.raises(NotImplementedError):
        car_shard_factorization(["ecog"] * 3, ["good"] * 3, ["average", "ch0", "average"], 0, 3)
