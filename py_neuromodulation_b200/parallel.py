"""Channel-sharded multi-GPU execution of the offline hot path.

After re-referencing every hot-path feature is computed per channel, so the recording shards by channel with no
halo: rank r of R owns channels [r*C/R, (r+1)*C/R) for ALL windows.  The data path has exactly two exchanges:

1. common-average reference: all-reduce(sum) of the per-sample channel-group sums (G x T float64), slice by slice on the
   library's side stream so that transfers, reductions and window kernels overlap -- ``nm_upload_begin_f32`` /
   ``nm_upload_slice_sums`` / ``nm_upload_slice_reduced`` / ``nm_upload_finish`` in the C ABI;
2. the (n_windows x F_local) float64 result blocks meet on rank 0.  On one node (the default) every rank copies the rows of
   each finished chunk over its OWN PCIe link straight into its column range of one page-locked POSIX shared-memory matrix
   (``nm_set_output_pitch`` + ``nm_host_register``), overlapped with the next chunk's kernels -- no collective, no funnel
   through rank 0's link.  ``shared_host=False`` (or ranks on different nodes) uses one NCCL gather to rank 0 instead.

The collectives go through ``torch.distributed`` (NCCL over NVLink on GPUs; gloo in the CPU tests, where the
"device" buffers of the thread-emulated library are host memory).  torch is plumbing only: no torch kernel touches
the samples.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def shard_bounds(n_channels: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced channel block of ``rank`` (first ``n % world`` ranks get one extra channel)."""
    base, extra = divmod(n_channels, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def car_shard_factorization(types: list[str], status: list[str], refs: list[str], lo: int, hi: int):
    """Factored re-reference (see ``csrc/nm_prep.cuh``) of the local channels [lo, hi) of a GLOBAL channel table.

    Supports "average" (per channel type, good channels only) and "None"; bipolar references would need the
    neighbour's samples on the same rank and are rejected.  Returns (n_groups, group_of, gcoef, sp_ptr, sp_col, sp_val).
    """
    n_loc = hi - lo
    kinds = sorted({t for t, s, r in zip(types, status, refs) if s == "good" and str(r).lower() == "average"})
    if len(kinds) > 8:
        raise ValueError("more than 8 channel types use an average reference")
    members = {k: [i for i, (t, s) in enumerate(zip(types, status)) if t == k and s == "good"] for k in kinds}
    group_of = np.full(n_loc, -1, dtype=np.int32)
    gcoef = np.zeros((n_loc, len(kinds)))
    diag = np.ones(n_loc)
    for j in range(n_loc):
        i = lo + j
        ref = str(refs[i]).lower()
        if status[i] == "good" and types[i] in kinds:
            group_of[j] = kinds.index(types[i])  # the channel contributes to its type's group sum
        if status[i] != "good" or ref == "none":
            continue
        if ref != "average":
            raise NotImplementedError("channel-sharded runs support 'average' and 'None' references only")
        n_ref = len(members[types[i]]) - 1  # all good channels of the type except i itself
        if n_ref <= 0:
            continue
        g = kinds.index(types[i])
        gcoef[j, g] = -1.0 / n_ref
        diag[j] = 1.0 + 1.0 / n_ref  # the group sum contains x_i itself
    sp_ptr = np.arange(n_loc + 1, dtype=np.int32)
    sp_col = np.arange(n_loc, dtype=np.int32)
    return len(kinds), group_of, gcoef, sp_ptr, sp_col, diag


def merge_permutation(settings, global_names: list[str], sfreq: float, window_samples: int, world: int):
    """Columns of the rank-major concatenation [rank 0 block | rank 1 block | ...] -> reference column order.

    Returns (reference_columns, perm) with ``reference_matrix = gathered[:, perm]``.  Plug-ins such as FFT / Welch
    order their keys band -> estimator -> channel (channel fastest), so the shards interleave in the reference order.
    """
    from .stream.data_processor import build_specs

    ref_cols = build_specs(settings, global_names, sfreq, window_samples)[2]
    concat: list[str] = []
    for r in range(world):
        lo, hi = shard_bounds(len(global_names), world, r)
        concat += build_specs(settings, global_names[lo:hi], sfreq, window_samples)[2]
    where = {k: i for i, k in enumerate(concat)}
    if len(where) != len(concat) or set(where) != set(ref_cols):
        raise ValueError("shard columns do not tile the reference columns")
    return ref_cols, np.array([where[k] for k in ref_cols], dtype=np.int64)


class _DeviceView:
    """Expose a raw device pointer through ``__cuda_array_interface__`` so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8") -> None:
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": (int(n),), "typestr": typestr, "version": 2,
                                         "strides": None}


def wrap_buffer(ptr: int, n: int, on_gpu: bool):
    """torch tensor over ``n`` float64 values at ``ptr`` (device memory on GPUs, host memory for the emulated library)."""
    import torch

    if on_gpu:
        return torch.as_tensor(_DeviceView(ptr, n), device="cuda")
    arr = np.ctypeslib.as_array((C.c_double * n).from_address(ptr))
    return torch.from_numpy(arr)


class ShardedRun:
    """One rank's part of a channel-sharded offline run around an already-built local :class:`Pipeline`."""

    def __init__(self, pipe, on_gpu: bool = True, shared_host: bool = True) -> None:
        self.pipe = pipe
        self.on_gpu = on_gpu
        self.shared_host = shared_host
        self._shm = None
        self._shm_key = None

    # ------------------------------------------------------------------ shared host matrix (single node)
    def _ensure_shared(self, n_windows: int):
        """(matrix view, first column of this rank) of the node-wide result matrix; created on first use."""
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory

        key = (n_windows, self.pipe.F)
        if self._shm_key == key:
            return self._shm_view, self._shm_col0
        self._release_shared()
        world, rank = dist.get_world_size(), dist.get_rank()
        widths = [None] * world
        dist.all_gather_object(widths, int(self.pipe.F))
        total = int(sum(widths))
        n_bytes = n_windows * total * 8
        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=max(n_bytes, 8))
            name[0] = shm.name
        dist.broadcast_object_list(name, src=0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name[0])
            try:  # attaching registers the segment with this process's resource tracker too (CPython < 3.13): only the owner unlinks
                from multiprocessing import resource_tracker

                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:
                pass
        view = np.ndarray((n_windows, total), dtype=np.float64, buffer=shm.buf)
        addr = view.ctypes.data
        _lib.check(self.pipe.lib.nm_host_register(C.c_void_p(addr), n_bytes))
        _lib.check(self.pipe.lib.nm_set_output_pitch(self.pipe._h, total))
        self._shm, self._shm_view, self._shm_addr = shm, view, addr
        self._shm_col0 = int(sum(widths[:rank]))
        self._shm_key = key
        self._shm_owner = rank == 0
        dist.barrier()
        return view, self._shm_col0

    def _release_shared(self) -> None:
        if self._shm is None:
            return
        try:
            self.pipe.synchronize()
            self.pipe.lib.nm_host_unregister(C.c_void_p(self._shm_addr))
            self.pipe.lib.nm_set_output_pitch(self.pipe._h, 0)
        except Exception:
            pass
        self._shm_view = None
        shm, self._shm = self._shm, None
        self._shm_key = None
        shm.close()
        if self._shm_owner:
            try:
                shm.unlink()
            except FileNotFoundError:
                pass

    def close(self) -> None:
        self._release_shared()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self._release_shared()
        except Exception:
            pass

    def _use_shared(self) -> bool:
        import torch.distributed as dist

        return self.shared_host and dist.is_initialized() and dist.get_world_size() > 1

    def upload(self, data_f32: np.ndarray) -> None:
        """Asynchronous, sliced: H2D of the local shard, per-slice group sums, per-slice all-reduce on the library's side stream.

        Nothing here blocks the host: the re-reference of a slice happens inside ``run`` right before the first chunk of
        windows that needs it, so transfers, reductions and window kernels of different slices overlap."""
        import contextlib

        import torch.distributed as dist

        p = self.pipe
        a = np.ascontiguousarray(data_f32, dtype=np.float32)
        _lib.check(p.lib.nm_upload_begin_f32(p._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[1]))
        p._keep_data = a
        n_slices, slice_len, n_groups, pitch = C.c_int(), C.c_longlong(), C.c_int(), C.c_longlong()
        _lib.check(p.lib.nm_upload_slices(p._h, C.byref(n_slices), C.byref(slice_len), C.byref(n_groups), C.byref(pitch)))
        ptr, n = C.c_void_p(), C.c_longlong()
        _lib.check(p.lib.nm_group_sums_device_ptr(p._h, C.byref(ptr), C.byref(n)))
        multi = dist.is_initialized() and dist.get_world_size() > 1
        ctx = contextlib.nullcontext()
        sums = None
        if multi:
            sums = wrap_buffer(ptr.value, n.value, self.on_gpu).view(n_groups.value, pitch.value)
            if self.on_gpu:
                import torch

                h = C.c_void_p()
                _lib.check(p.lib.nm_side_stream_handle(p._h, C.byref(h)))
                ctx = torch.cuda.stream(torch.cuda.ExternalStream(h.value))  # the collective is ordered on the library's stream
        T = a.shape[1]
        for k in range(n_slices.value):
            _lib.check(p.lib.nm_upload_slice_sums(p._h, k))
            if multi:
                t0, t1 = k * slice_len.value, min(T, (k + 1) * slice_len.value)
                with ctx:
                    for g in range(n_groups.value):
                        dist.all_reduce(sums[g, t0:t1], op=dist.ReduceOp.SUM)
            _lib.check(p.lib.nm_upload_slice_reduced(p._h, k))
        _lib.check(p.lib.nm_upload_finish(p._h))

    def run(self, starts: np.ndarray) -> None:
        p = self.pipe
        if not self._use_shared():
            p.run(starts, download=False)
            p.synchronize()
            return
        s = np.ascontiguousarray(starts, dtype=np.int64)
        view, col0 = self._ensure_shared(int(s.size))
        dst = view.ctypes.data + col0 * 8
        _lib.check(p.lib.nm_run_windows(p._h, s.ctypes.data_as(C.POINTER(C.c_longlong)), int(s.size), C.c_void_p(dst)))

    def gather(self, n_windows: int):
        """Gather the (n_windows, F_local) blocks to rank 0; returns a host array (n_windows, sum F_local) there, else None.

        Device staging blocks and the pinned host matrix are allocated once and reused (the returned array is a view
        of that pinned buffer: copy it if it must survive the next call).
        """
        import torch
        import torch.distributed as dist

        if self._use_shared() and self._shm_key == (n_windows, self.pipe.F):
            dist.barrier()  # every rank has returned from run(): its block is in the shared matrix
            return self._shm_view if dist.get_rank() == 0 else None
        ptr, rows, cols = self.pipe.result_device_ptr()
        local = wrap_buffer(ptr, rows * cols, self.on_gpu)[: n_windows * cols].view(n_windows, cols)
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return local.cpu().numpy().copy() if self.on_gpu else local.numpy().copy()
        world, rank = dist.get_world_size(), dist.get_rank()
        key = (n_windows, cols)
        if getattr(self, "_gather_key", None) != key:
            # shards may differ by one channel: agree on the widest block once, pad, gather, trim
            widths = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
            dist.all_gather(widths, torch.tensor([cols], dtype=torch.int64, device=local.device))
            self._widths = [int(w.item()) for w in widths]
            wmax = max(self._widths)
            self._block = None if cols == wmax else local.new_zeros(n_windows, wmax)
            self._parts = [local.new_empty(n_windows, wmax) for _ in range(world)] if rank == 0 else None
            self._host = None
            self._full = None
            if rank == 0:
                self._host = torch.empty((n_windows, sum(self._widths)), dtype=local.dtype, pin_memory=self.on_gpu)
            self._gather_key = key
        if self._block is None:
            block = local if local.is_contiguous() else local.contiguous()
        else:
            self._block[:, :cols].copy_(local)
            block = self._block
        if rank != 0:
            dist.gather(block, gather_list=None, dst=0)
            return None
        dist.gather(block, gather_list=self._parts, dst=0)
        # concatenate the column blocks ON THE DEVICE (one pass at HBM speed), then a single contiguous D2H into the
        # pinned matrix: a strided device->host copy per block goes through a slow path (87 ms instead of 3 ms for 147 MB)
        if self._full is None:
            self._full = local.new_empty(n_windows, sum(self._widths))
        torch.cat([part[:, :w] for part, w in zip(self._parts, self._widths)], dim=1, out=self._full)
        self._host.copy_(self._full, non_blocking=True)
        if self.on_gpu:
            torch.cuda.current_stream().synchronize()
        return self._host.numpy()
