"""Per-window latency of the streaming entry, default settings (all default plug-ins + normaliser).

    python tools/stream_latency.py [n_channels ...]

Per channel count: DataProcessor.process (dict), nm_process_window (pageable buffers in and out), the slot API (window written
into the page-locked slot, submit, wait) with eager launches and with CUDA-graph replay, and the pipelined slot API (window g + 1
submitted before window g is collected: sustained windows/s).
"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import py_neuromodulation_b200 as nm  # noqa: E402
from py_neuromodulation_b200.stream.window_stream import WindowStream  # noqa: E402
from py_neuromodulation_b200.utils.channels import get_default_channels_from_data  # noqa: E402

N_WIN = 400


def med(ts):
    a = np.array(ts[50:]) * 1e3
    return f"{np.median(a):.3f} ms (p95 {np.percentile(a, 95):.3f})"


def main():
    chans = [int(a) for a in sys.argv[1:]] or [8, 64, 256]
    for n_ch in chans:
        x = np.random.default_rng(0).random((n_ch, 1000 + 100 * N_WIN))
        s = nm.NMSettings.get_default()
        ch = get_default_channels_from_data(x)
        wins = [np.ascontiguousarray(x[:, 100 * k : 100 * k + 1000]) for k in range(N_WIN)]

        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
        t_dict = []
        for w in wins:
            t0 = time.perf_counter()
            dp.process(w)
            t_dict.append(time.perf_counter() - t0)
        F = dp.plan(1000).pipe.F

        dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
        pipe = dp.plan(1000).pipe
        t_raw = []
        for w in wins:
            t0 = time.perf_counter()
            pipe.process_window(w)
            t_raw.append(time.perf_counter() - t0)

        res = {}
        for graph in (False, True):
            dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
            ws = WindowStream(dp, 1000, slots=2, graph=graph)
            t_slot, t_gpu = [], []
            for w in wins:
                t0 = time.perf_counter()
                ws.next_input()[...] = w
                t1 = time.perf_counter()
                ws.submit()
                ws.collect()
                t2 = time.perf_counter()
                t_slot.append(t2 - t0)
                t_gpu.append(t2 - t1)
            stats = ws.pipe.stream_stats()
            ws.close()
            # pipelined: keep one window in flight while the host fills the next slot
            dp = nm.DataProcessor(sfreq=1000, settings=s, channels=ch, line_noise=50, verbose=False)
            ws = WindowStream(dp, 1000, slots=2, graph=graph)
            for w in wins[:50]:
                ws.process(w)
            t0 = time.perf_counter()
            for w in wins[50:]:
                ws.next_input()[...] = w
                ws.submit()
                if ws.in_flight == ws.slots:
                    ws.collect()
            ws.drain()
            rate = (N_WIN - 50) / (time.perf_counter() - t0)
            ws.close()
            res[graph] = (med(t_slot), med(t_gpu), rate, stats)
        print(f"{n_ch:4d} ch, F = {F}:")
        print(f"    DataProcessor.process (dict)          {med(t_dict)}")
        print(f"    nm_process_window (pageable in/out)   {med(t_raw)}")
        for graph in (False, True):
            a, b, rate, stats = res[graph]
            print(f"    slot API, {'graph replay' if graph else 'eager       '}: copy+submit+wait {a}; submit+wait only {b}; pipelined {rate:.0f} windows/s; {stats}")


if __name__ == "__main__":
    main()
