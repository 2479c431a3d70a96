// nm_stream.cuh -- entry points of the streaming path (state and design notes: nm_stream.h).  Included at the end of nm_pipeline.cu.
//
//   nm_stream_open(p, n_slots, input_f32, use_graph)   page-locked slot ring, parameter blocks, device geometry of ONE window
//   nm_stream_input(p, slot, &ptr)                     where the producer writes the (n_raw_rows x window) samples of the next window
//   nm_stream_submit(p, slot)                          enqueue the window (H2D, all kernels, D2H of the feature row); does not block
//   nm_stream_wait(p, slot, &features)                 block until that window's feature row is in the slot's output block
//   nm_stream_close(p)
// nm_process_window (caller-owned buffers, synchronous) is a thin layer over these.
#pragma once

static void nm_stream_release(nm_pipeline* p) {
    if (!p || !p->strm) return;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    for (auto& s : p->strm->slots) {
#ifndef NM_EMULATE
        if (s.gs.exec) cudaGraphExecDestroy(s.gs.exec);
        if (s.graph) cudaGraphDestroy(s.graph);
#endif
        if (s.in_host) cudaFreeHost(s.in_host);
        if (s.out_host) cudaFreeHost(s.out_host);
        if (s.par_host) cudaFreeHost(s.par_host);
        if (s.done) cudaEventDestroy(s.done);
    }
    p->strm.reset();
}

extern "C" int nm_stream_close(nm_pipeline* p) {
    NM_P_CHECK(p);
    nm_stream_release(p);
    return 0;
}

// use_graph: 1 replay a captured CUDA graph per window, 0 launch eagerly, -1 environment (NMB200_STREAM_GRAPH, default 1)
extern "C" int nm_stream_open(nm_pipeline* p, int n_slots, int input_f32, int use_graph) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized, "call nm_finalize first");
    NM_CHECK(n_slots >= 1 && n_slots <= 64, "n_slots must be in [1, 64]");
    cudaSetDevice(p->device);
    nm_upload_join(p);  // (a deferred batched upload may still be staging)
    nm_stream_release(p);
    NM_CUDA_CHECK(cudaStreamSynchronize(p->copy_stream));
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    auto st = std::make_unique<NmStream>();
    st->n_slots = n_slots;
    st->f32 = input_f32 != 0;
    if (use_graph < 0) {
        const char* env = getenv("NMB200_STREAM_GRAPH");
        use_graph = env ? atoi(env) : 1;
    }
#ifdef NM_EMULATE
    use_graph = 0;
#endif
    st->use_graph = use_graph != 0 && !p->profiling;
    const size_t esz = st->f32 ? 4 : 8;
    st->slots.resize(n_slots);
    for (auto& s : st->slots) {
        NM_CUDA_CHECK(cudaMallocHost(&s.in_host, (size_t)p->C_all * p->Win * esz));
        NM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&s.out_host), (size_t)p->F * sizeof(double)));
        NM_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&s.par_host), NM_STREAM_PAR_BYTES));
        NM_CUDA_CHECK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    }
    if (st->d_par.ensure(NM_STREAM_PAR_BYTES)) return -1;
    // device geometry of a one-window "recording" (what nm_upload_impl sets up for every call of the old entry)
    p->n_slices = 0;
    p->slices_prepped = 0;
    p->raw_f64 = !st->f32;
    p->T = p->Win;
    p->raw_pitch = (p->Win + 3) & ~3LL;
    p->xr_pitch = (p->Win + 1) & ~1LL;
    p->nanblk_pitch = (p->Win + 31) / 32;
    if (nm_stage_buffers(p, esz)) return -1;
    const long long zero = 0;
    if (p->d_starts.upload(&zero, 1, p->stream)) return -1;
    if (p->d_out.ensure((size_t)p->F * sizeof(double))) return -1;
    if (p->has_nan_cols && p->d_nanflags.ensure((size_t)p->C_all)) return -1;
    if (nm_fused_spec_upload(p)) return -1;
    if (p->norm && p->norm->n_cols) {
        if (p->norm->d_ext.ensure((size_t)std::max(1, p->norm->n_keep) * p->norm->n_cols * sizeof(double))) return -1;
    }
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->out_rows = 1;
    p->have_data = true;
    p->upload_pending = false;
    p->resident_uses_gsum = false;
    p->strm = std::move(st);
    return 0;
}

extern "C" int nm_stream_input(nm_pipeline* p, int slot, void** ptr, long long* bytes) {
    NM_P_CHECK(p);
    NM_CHECK(p->strm && slot >= 0 && slot < p->strm->n_slots, "no such streaming slot (call nm_stream_open)");
    if (ptr) *ptr = p->strm->slots[slot].in_host;
    if (bytes) *bytes = (long long)p->C_all * p->Win * (p->strm->f32 ? 4 : 8);
    return 0;
}

// everything one window enqueues, in stream order; runs eagerly, under stream capture, or as the patch pass of a graph replay
static int nm_stream_pass(nm_pipeline* p, int slot) {
    NmStream& st = *p->strm;
    NmStreamSlot& s = st.slots[slot];
    const size_t esz = st.f32 ? 4 : 8;
    st.cur = slot;
    st.par_used = 0;
    int rc = 0;
    do {
        // H2D of the window: slot rows are dense (Win samples), device rows have the padded pitch
        if (p->raw_pitch == p->Win) {
            if (!nm_gs_updating() && cudaMemcpyAsync(p->d_raw.p, s.in_host, (size_t)p->C_all * p->Win * esz, cudaMemcpyHostToDevice, p->stream) != cudaSuccess) { rc = -1; break; }
        } else {
            if (!nm_gs_updating() && cudaMemcpy2DAsync(p->d_raw.p, (size_t)p->raw_pitch * esz, s.in_host, (size_t)p->Win * esz, (size_t)p->Win * esz,
                                                       (size_t)p->C_all, cudaMemcpyHostToDevice, p->stream) != cudaSuccess) { rc = -1; break; }
        }
        nm_launch_prep(p, nm_prep_args(p));
        if (!nm_gs_updating() && cudaMemsetAsync(p->d_out.p, 0, (size_t)p->F * sizeof(double), p->stream) != cudaSuccess) { rc = -1; break; }
        if (p->bursts && p->bursts->prepare(p, 1)) { rc = -1; break; }
        if (nm_run_chunk(p, 0, 1)) { rc = -1; break; }
        if (p->norm && p->norm->run(p, 1)) { rc = -1; break; }
        if (p->has_nan_cols) nm_nan_fill(p, 0, 1);
        if (!nm_gs_updating() && cudaMemcpyAsync(s.out_host, p->d_out.p, (size_t)p->F * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) { rc = -1; break; }
    } while (false);
    st.cur = -1;
    if (rc) nm_set_error("streaming pass failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

extern "C" int nm_stream_submit(nm_pipeline* p, int slot) {
    NM_P_CHECK(p);
    NM_CHECK(p->strm && slot >= 0 && slot < p->strm->n_slots, "no such streaming slot (call nm_stream_open)");
    NmStream& st = *p->strm;
    NmStreamSlot& s = st.slots[slot];
    NM_CHECK(!s.busy, "slot %d still holds an un-collected window (nm_stream_wait)", slot);
    NM_CHECK(p->T == p->Win && p->n_slices == 0 && p->raw_f64 == !st.f32, "a recording was uploaded after nm_stream_open: re-open the stream");
    cudaSetDevice(p->device);
    int rc = 0;
#ifndef NM_EMULATE
    if (st.use_graph && s.gs.exec && !s.gs.broken) {
        // replay: the host pass only patches the arguments that changed since the last window of this slot
        s.gs.mode = 2;
        s.gs.cursor = 0;
        nm_graph_session = &s.gs;
        rc = nm_stream_pass(p, slot);
        nm_graph_session = nullptr;
        s.gs.mode = 0;
        NM_CHECK(rc == 0 && !s.gs.broken && s.gs.cursor == s.gs.nodes.size(),
                 "the streaming launch sequence changed under a captured graph (%zu of %zu kernels): re-open the stream", s.gs.cursor,
                 s.gs.nodes.size());
        NM_CUDA_CHECK(cudaGraphLaunch(s.gs.exec, p->stream));
        st.graph_launches++;
    } else if (st.use_graph && s.warm && !s.gs.broken) {
        // second window of the slot: every buffer exists, capture the sequence (forked side streams join before the end)
        s.gs.nodes.clear();
        s.gs.mode = 1;
        nm_graph_session = &s.gs;
        NM_CUDA_CHECK(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeRelaxed));
        rc = nm_stream_pass(p, slot);
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(p->stream, &g);
        nm_graph_session = nullptr;
        s.gs.mode = 0;
        NM_CHECK(rc == 0 && ce == cudaSuccess && g != nullptr, "stream capture of the window failed: %s", cudaGetErrorString(ce));
        if (s.gs.broken) {  // could not identify the kernel nodes: stay eager (the captured work still has to run once)
            cudaGraphExec_t once = nullptr;
            NM_CUDA_CHECK(cudaGraphInstantiate(&once, g, 0));
            NM_CUDA_CHECK(cudaGraphLaunch(once, p->stream));
            NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
            cudaGraphExecDestroy(once);
            cudaGraphDestroy(g);
        } else {
            s.graph = g;
            NM_CUDA_CHECK(cudaGraphInstantiate(&s.gs.exec, g, 0));
            NM_CUDA_CHECK(cudaGraphLaunch(s.gs.exec, p->stream));
            st.graph_launches++;
        }
    } else
#endif
    {
        rc = nm_stream_pass(p, slot);
        if (rc) return -1;
        s.warm = true;
    }
    NM_CUDA_CHECK(cudaEventRecord(s.done, p->stream));
    s.busy = true;
    st.windows++;
    return 0;
}

extern "C" int nm_stream_wait(nm_pipeline* p, int slot, const double** features) {
    NM_P_CHECK(p);
    NM_CHECK(p->strm && slot >= 0 && slot < p->strm->n_slots, "no such streaming slot (call nm_stream_open)");
    NmStreamSlot& s = p->strm->slots[slot];
    NM_CHECK(s.busy, "slot %d has no window in flight", slot);
    NM_CUDA_CHECK(cudaEventSynchronize(s.done));
    s.busy = false;
    if (features) *features = s.out_host;
    return 0;
}

// windows submitted, graph replays, kernel nodes per graph, argument patches applied so far
extern "C" int nm_stream_stats(nm_pipeline* p, long long* windows, long long* graph_launches, long long* graph_nodes, long long* patched) {
    NM_P_CHECK(p);
    NM_CHECK(p->strm, "no stream open");
    long long nodes = 0, pat = 0;
#ifndef NM_EMULATE
    for (auto& s : p->strm->slots) {
        nodes = std::max<long long>(nodes, (long long)s.gs.nodes.size());
        pat += s.gs.patched;
    }
#endif
    if (windows) *windows = p->strm->windows;
    if (graph_launches) *graph_launches = p->strm->graph_launches;
    if (graph_nodes) *graph_nodes = nodes;
    if (patched) *patched = pat;
    return 0;
}
