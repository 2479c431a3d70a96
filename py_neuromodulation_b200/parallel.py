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


class NativeComm:
    """NCCL communicator owned by libnmb200 (``nm_comm_*``, csrc/nm_comm.cuh): the collectives of the sharded path run inside the
    library, on its own streams -- no torch on the data path.

    ``NativeComm.from_env()`` is the rendezvous for ``torchrun``-style launches (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
    MASTER_PORT): rank 0 creates the 128-byte unique id and publishes it through a ``torch.distributed.TCPStore`` -- plumbing
    only.  Any other channel (MPI, a file) works with ``NativeComm(id_bytes, rank, world, device)``."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int) -> None:
        self.lib = _lib.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.nm_comm_create(buf, self.rank, self.world, self.device, C.byref(h)))
        self._h = h
        self.store = None

    @staticmethod
    def new_unique_id() -> bytes:
        lib = _lib.load()
        buf = (C.c_ubyte * 128)()
        _lib.check(lib.nm_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_env(cls, device: int | None = None, port_offset: int = 17) -> "NativeComm":
        import os
        from datetime import timedelta

        from torch.distributed import TCPStore

        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        host = os.environ.get("MASTER_ADDR", "127.0.0.1")
        port = int(os.environ.get("MASTER_PORT", "29500")) + port_offset  # next to (not on) torchrun's own store
        store = TCPStore(host, port, world, is_master=(rank == 0), timeout=timedelta(seconds=120), wait_for_workers=False)
        if rank == 0:
            store.set("nmb200_nccl_id", cls.new_unique_id())
        uid = bytes(store.get("nmb200_nccl_id"))
        comm = cls(uid, rank, world, device)
        comm.store = store  # keeps the rendezvous alive; also carries small host-side objects (column widths, shm names)
        return comm

    def barrier(self) -> None:
        _lib.check(self.lib.nm_comm_barrier(self._h))

    def max(self, value: float) -> float:
        v = C.c_double(float(value))
        _lib.check(self.lib.nm_comm_allreduce_max(self._h, C.byref(v)))
        return float(v.value)

    @property
    def collectives(self) -> int:
        return int(self.lib.nm_comm_collectives(self._h))

    def exchange(self, key: str, value) -> list:
        """All-gather of a small picklable host object through the rendezvous store (set-up time only)."""
        import pickle

        assert self.store is not None, "NativeComm.exchange needs the rendezvous store (from_env)"
        self._xchg = getattr(self, "_xchg", 0) + 1
        tag = f"{key}/{self._xchg}"
        self.store.set(f"{tag}/{self.rank}", pickle.dumps(value))
        return [pickle.loads(self.store.get(f"{tag}/{r}")) for r in range(self.world)]

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.nm_comm_destroy(self._h)
            self._h = None

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class ShardedRun:
    """One rank's part of a channel-sharded offline run around an already-built local :class:`Pipeline`.

    ``comm=None``: collectives through ``torch.distributed`` (gloo in the CPU tests, NCCL on GPUs), driven slice by slice from the
    host.  ``comm=NativeComm``: the library issues them itself (``nm_upload_sharded_f32`` / ``nm_gather_results``)."""

    def __init__(self, pipe, on_gpu: bool = True, shared_host: bool = True, comm: "NativeComm | None" = None) -> None:
        self.pipe = pipe
        self.on_gpu = on_gpu
        self.shared_host = shared_host
        self.comm = comm
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
        world, rank = self._world_rank()
        if self.comm is not None:
            widths = self.comm.exchange("widths", int(self.pipe.F))
        else:
            widths = [None] * world
            dist.all_gather_object(widths, int(self.pipe.F))
        total = int(sum(widths))
        n_bytes = n_windows * total * 8
        name = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=max(n_bytes, 8))
            name[0] = shm.name
        if self.comm is not None:
            name[0] = self.comm.exchange("shm", name[0])[0]
        else:
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
        self._widths_all = [int(w) for w in widths]
        self._barrier()
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

    def _world_rank(self) -> tuple[int, int]:
        if self.comm is not None:
            return self.comm.world, self.comm.rank
        import torch.distributed as dist

        return (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)

    def _barrier(self) -> None:
        if self.comm is not None:
            self.comm.barrier()
        else:
            import torch.distributed as dist

            dist.barrier()

    def _use_shared(self) -> bool:
        return self.shared_host and self._world_rank()[0] > 1

    def upload(self, data_f32: np.ndarray) -> None:
        """Asynchronous, sliced: H2D of the local shard, per-slice group sums, per-slice all-reduce on the library's side stream.

        Nothing here blocks the host: the re-reference of a slice happens inside ``run`` right before the first chunk of
        windows that needs it, so transfers, reductions and window kernels of different slices overlap."""
        p = self.pipe
        a = np.ascontiguousarray(data_f32, dtype=np.float32)
        if self.comm is not None:  # sums + ncclAllReduce per slice are enqueued by the library itself (no torch on this path)
            _lib.check(p.lib.nm_upload_sharded_f32(p._h, self.comm._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[1]))
            p._keep_data = a
            return
        import contextlib

        import torch.distributed as dist

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
            self._barrier()  # every rank has returned from run(): its block is in the shared matrix
            return self._shm_view if self._world_rank()[1] == 0 else None
        if self.comm is not None:  # one grouped ncclSend / ncclRecv gather inside the library
            if getattr(self, "_native_key", None) != (n_windows, self.pipe.F):
                self._widths_native = [int(w) for w in self.comm.exchange("gw", int(self.pipe.F))]
                self._native_host = None
                if self.comm.rank == 0:
                    from . import _lib as L

                    n_bytes = n_windows * sum(self._widths_native) * 8
                    ptr = C.c_void_p()
                    L.check(self.pipe.lib.nm_host_alloc(C.byref(ptr), n_bytes))
                    buf = (C.c_char * n_bytes).from_address(ptr.value)
                    self._native_host = np.frombuffer(buf, dtype=np.float64).reshape(n_windows, sum(self._widths_native))
                self._native_key = (n_windows, self.pipe.F)
            w = (C.c_int * self.comm.world)(*self._widths_native)
            dst = self._native_host.ctypes.data_as(C.c_void_p) if self.comm.rank == 0 else None
            _lib.check(self.pipe.lib.nm_gather_results(self.pipe._h, self.comm._h, int(n_windows), w, dst))
            return self._native_host
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
