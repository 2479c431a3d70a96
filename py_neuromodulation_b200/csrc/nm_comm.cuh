// nm_comm.cuh -- the collectives of the channel-sharded path inside the library: a maintainer who binds
// include/nmb200.h gets the multi-GPU path without torch.
//
// One process per GPU.  The application distributes a 128-byte NCCL unique id (rank 0: nm_comm_unique_id) by whatever
// rendezvous it has (MPI, a file, torch's TCP store, ...); every rank then calls nm_comm_create.  The data path has exactly two
// exchanges, both enqueued by the library on its own streams:
//   nm_upload_sharded_f32   H2D of the local channel shard in time slices; per slice: local per-sample group sums
//                           (nm_gsum_kernel) followed by ncclAllReduce(sum, float64) of that slice ON THE SAME (reduction) stream;
//                           the window kernels of nm_run_windows wait per slice on an event, so transfers, reductions and window
//                           kernels of different slices overlap and the call returns without blocking the host
//   nm_gather_results       the (n_windows x F_r) result blocks -> rank 0 (grouped ncclSend / ncclRecv; shards may be uneven),
//                           concatenated into rank-major column order by one kernel, one contiguous D2H
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy the process already has -- e.g. torch's -- or the system one), so
// libnmb200.so keeps loading on machines without NCCL; nm_comm_* then fail with an error text.  The thread-emulated TEST build has
// no NCCL: the CPU tests cover the same protocol through the host-driven entry points (nm_upload_begin_f32, ...) over gloo.
#pragma once

#ifndef NM_EMULATE
#include <dlfcn.h>
#include <nccl.h>

struct NmNccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

static NmNccl& nm_nccl() {
    static NmNccl n;
    if (n.handle || n.ok) return n;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.handle) break;
    }
    if (!n.handle) return n;
#define NM_NCCL_SYM(f) n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.handle, "nccl" #f))
    NM_NCCL_SYM(GetUniqueId); NM_NCCL_SYM(CommInitRank); NM_NCCL_SYM(CommDestroy); NM_NCCL_SYM(AllReduce); NM_NCCL_SYM(Send);
    NM_NCCL_SYM(Recv); NM_NCCL_SYM(GroupStart); NM_NCCL_SYM(GroupEnd); NM_NCCL_SYM(GetErrorString); NM_NCCL_SYM(GetVersion);
#undef NM_NCCL_SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.Send && n.Recv && n.GroupStart && n.GroupEnd && n.GetErrorString;
    return n;
}

#define NM_NCCL_CHECK(expr)                                                                                       \
    do {                                                                                                          \
        ncclResult_t _r = (expr);                                                                                 \
        if (_r != ncclSuccess) {                                                                                  \
            nm_set_error("%s failed: %s (%s:%d)", #expr, nm_nccl().GetErrorString(_r), __FILE__, __LINE__);       \
            return -1;                                                                                            \
        }                                                                                                         \
    } while (0)

struct nm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;  // small control collectives (barrier, max)
    DevBuf d_scalar, d_parts, d_full;
    long long collectives = 0;
};

// concat[w, col0[r] + j] = parts[r][w, j]: rank-major column blocks of the gathered result (one pass at HBM speed)
NM_GLOBAL void nm_concat_kernel(const double* parts, const long long* part_off, const int* widths, const int* col0, int world, int n_windows,
                                int total, double* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_windows * total) return;
    const int w = (int)(idx / total), c = (int)(idx - (long long)w * total);
    int r = 0;
    while (r + 1 < world && c >= col0[r + 1]) ++r;
    out[idx] = parts[part_off[r] + (long long)w * widths[r] + (c - col0[r])];
}

extern "C" int nm_comm_unique_id(unsigned char* id128) {
    NM_CHECK(id128, "id128 is NULL");
    NM_CHECK(nm_nccl().ok, "NCCL is not available (dlopen libnccl.so.2 failed)");
    ncclUniqueId id;
    NM_NCCL_CHECK(nm_nccl().GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return 0;
}

extern "C" int nm_comm_create(const unsigned char* id128, int rank, int world, int device, nm_comm** out) {
    NM_CHECK(id128 && out, "NULL argument");
    NM_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    NM_CHECK(nm_nccl().ok, "NCCL is not available (dlopen libnccl.so.2 failed)");
    NM_CUDA_CHECK(cudaSetDevice(device));
    auto c = std::make_unique<nm_comm>();
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NM_NCCL_CHECK(nm_nccl().CommInitRank(&c->comm, world, id, rank));
    NM_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    if (c->d_scalar.ensure(64)) return -1;
    *out = c.release();
    return 0;
}

extern "C" void nm_comm_destroy(nm_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->comm) nm_nccl().CommDestroy(c->comm);
    delete c;
}

extern "C" int nm_comm_rank(const nm_comm* c) { return c ? c->rank : -1; }
extern "C" int nm_comm_size(const nm_comm* c) { return c ? c->world : 0; }
extern "C" long long nm_comm_collectives(const nm_comm* c) { return c ? c->collectives : 0; }

// max over ranks of one double (timing: "time on the device, max over ranks"); doubles as a barrier
extern "C" int nm_comm_allreduce_max(nm_comm* c, double* value) {
    NM_CHECK(c && value, "NULL argument");
    NM_CUDA_CHECK(cudaSetDevice(c->device));
    NM_CUDA_CHECK(cudaMemcpyAsync(c->d_scalar.p, value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NM_NCCL_CHECK(nm_nccl().AllReduce(c->d_scalar.p, c->d_scalar.p, 1, ncclDouble, ncclMax, c->comm, c->stream));
    NM_CUDA_CHECK(cudaMemcpyAsync(value, c->d_scalar.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NM_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->collectives++;
    return 0;
}
extern "C" int nm_comm_barrier(nm_comm* c) {
    double v = 0.0;
    return nm_comm_allreduce_max(c, &v);
}

extern "C" int nm_upload_sharded_f32(nm_pipeline* p, nm_comm* c, const float* data, long long n_samples, long long pitch) {
    NM_P_CHECK(p);
    NM_CHECK(c, "comm is NULL");
    if (nm_upload_begin_f32(p, data, n_samples, pitch)) return -1;
    for (int k = 0; k < p->n_slices; ++k) {
        if (nm_upload_slice_sums(p, k)) return -1;
        if (c->world > 1) {
            const long long t0 = (long long)k * p->slice_len, len = std::min<long long>(p->slice_len, p->T - t0);
            // one grouped launch for all reference groups of the slice, ordered behind the sums on the reduction stream
            NM_NCCL_CHECK(nm_nccl().GroupStart());
            for (int g = 0; g < p->G; ++g) {
                double* s = p->d_gsum.as<double>() + (size_t)g * p->gsum_pitch + t0;
                NM_NCCL_CHECK(nm_nccl().AllReduce(s, s, (size_t)len, ncclDouble, ncclSum, c->comm, p->red_stream));
            }
            NM_NCCL_CHECK(nm_nccl().GroupEnd());
            c->collectives++;
        }
        if (nm_upload_slice_reduced(p, k)) return -1;
    }
    return nm_upload_finish(p);
}

// resident recording of a sharded run: recompute the local group sums, all-reduce them and redo the window-independent
// preprocessing -- the sharded counterpart of nm_prepare_resident (bench: `value` of an N-GPU step contains the exchange)
extern "C" int nm_prepare_resident_sharded(nm_pipeline* p, nm_comm* c) {
    NM_P_CHECK(p);
    NM_CHECK(c, "comm is NULL");
    NM_CHECK(p->have_data && !p->upload_pending && p->resident_uses_gsum, "no resident sharded recording (nm_upload_sharded_f32 first)");
    cudaSetDevice(p->device);
    if (nm_ensure_prepped(p, p->T)) return -1;
    // the reduction stream must not overwrite the sums while window kernels of the previous run still read them
    NM_CUDA_CHECK(cudaEventRecord(p->ev_sync, p->stream));
    NM_CUDA_CHECK(cudaStreamWaitEvent(p->red_stream, p->ev_sync, 0));
    NmPrepArgs a = nm_prep_args(p);
    a.gsum_pitch = p->gsum_pitch;
    const unsigned grid = (unsigned)((p->T + 31) / 32);
    NM_LAUNCH(nm_gsum_kernel, dim3(grid), dim3(NM_PREP_THREADS), nm_prep_smem_bytes(), p->red_stream, a, p->d_gsum.as<double>());
    p->launches++;
    if (c->world > 1) {
        NM_NCCL_CHECK(nm_nccl().GroupStart());
        for (int g = 0; g < p->G; ++g) {
            double* s = p->d_gsum.as<double>() + (size_t)g * p->gsum_pitch;
            NM_NCCL_CHECK(nm_nccl().AllReduce(s, s, (size_t)p->T, ncclDouble, ncclSum, c->comm, p->red_stream));
        }
        NM_NCCL_CHECK(nm_nccl().GroupEnd());
        c->collectives++;
    }
    NM_CUDA_CHECK(cudaEventRecord(p->red_ev[0], p->red_stream));
    NM_CUDA_CHECK(cudaStreamWaitEvent(p->stream, p->red_ev[0], 0));
    return nm_prepare_resident(p);
}

extern "C" int nm_gather_results(nm_pipeline* p, nm_comm* c, int n_windows, const int* widths, double* out_host) {
    NM_P_CHECK(p);
    NM_CHECK(c && widths && n_windows > 0 && n_windows <= p->out_rows, "bad arguments");
    NM_CHECK(widths[c->rank] == p->F, "widths[%d] = %d but this pipeline has %d feature columns", c->rank, widths[c->rank], p->F);
    NM_CHECK(c->rank != 0 || out_host, "rank 0 needs the destination matrix");
    NM_CUDA_CHECK(cudaSetDevice(p->device));
    cudaStream_t s = p->stream;  // ordered behind the window kernels of this rank
    if (c->world == 1) return nm_download(p, out_host, n_windows);
    if (c->rank != 0) {
        NM_NCCL_CHECK(nm_nccl().Send(p->d_out.p, (size_t)n_windows * p->F, ncclDouble, 0, c->comm, s));
        c->collectives++;
        return 0;
    }
    std::vector<long long> off(c->world);
    std::vector<int> col0(c->world);
    long long tot_vals = 0;
    int total = 0;
    for (int r = 0; r < c->world; ++r) {
        off[r] = tot_vals; col0[r] = total;
        tot_vals += (long long)n_windows * widths[r];
        total += widths[r];
    }
    const size_t meta = (size_t)c->world * (sizeof(long long) + 2 * sizeof(int));
    if (c->d_parts.ensure((size_t)tot_vals * sizeof(double) + meta + 64) || c->d_full.ensure((size_t)tot_vals * sizeof(double))) return -1;
    double* parts = c->d_parts.as<double>();
    long long* d_off = reinterpret_cast<long long*>(parts + tot_vals);
    int* d_w = reinterpret_cast<int*>(d_off + c->world);
    int* d_c0 = d_w + c->world;
    NM_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), c->world * sizeof(long long), cudaMemcpyHostToDevice, s));
    NM_CUDA_CHECK(cudaMemcpyAsync(d_w, widths, c->world * sizeof(int), cudaMemcpyHostToDevice, s));
    NM_CUDA_CHECK(cudaMemcpyAsync(d_c0, col0.data(), c->world * sizeof(int), cudaMemcpyHostToDevice, s));
    NM_CUDA_CHECK(cudaStreamSynchronize(s));  // `off` / `col0` are temporaries
    NM_CUDA_CHECK(cudaMemcpyAsync(parts, p->d_out.p, (size_t)n_windows * p->F * sizeof(double), cudaMemcpyDeviceToDevice, s));
    NM_NCCL_CHECK(nm_nccl().GroupStart());
    for (int r = 1; r < c->world; ++r)
        NM_NCCL_CHECK(nm_nccl().Recv(parts + off[r], (size_t)n_windows * widths[r], ncclDouble, r, c->comm, s));
    NM_NCCL_CHECK(nm_nccl().GroupEnd());
    c->collectives++;
    const long long n_out = (long long)n_windows * total;
    NM_LAUNCH(nm_concat_kernel, dim3((unsigned)((n_out + 255) / 256)), dim3(256), 0, s, (const double*)parts, (const long long*)d_off,
              (const int*)d_w, (const int*)d_c0, c->world, n_windows, total, c->d_full.as<double>());
    p->launches++;
    NM_CUDA_CHECK(cudaMemcpyAsync(out_host, c->d_full.p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, s));
    NM_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

#else  // NM_EMULATE: the test build has no NCCL; the CPU tests drive the host-side protocol (nm_upload_begin_f32 ...) over gloo
struct nm_comm { int rank, world; };
extern "C" int nm_comm_unique_id(unsigned char*) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" int nm_comm_create(const unsigned char*, int, int, int, nm_comm**) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" void nm_comm_destroy(nm_comm*) {}
extern "C" int nm_comm_rank(const nm_comm*) { return -1; }
extern "C" int nm_comm_size(const nm_comm*) { return 0; }
extern "C" long long nm_comm_collectives(const nm_comm*) { return 0; }
extern "C" int nm_comm_allreduce_max(nm_comm*, double*) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" int nm_comm_barrier(nm_comm*) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" int nm_upload_sharded_f32(nm_pipeline*, nm_comm*, const float*, long long, long long) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" int nm_gather_results(nm_pipeline*, nm_comm*, int, const int*, double*) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
extern "C" int nm_prepare_resident_sharded(nm_pipeline*, nm_comm*) { nm_set_error("the thread-emulated test build has no NCCL"); return -1; }
#endif
