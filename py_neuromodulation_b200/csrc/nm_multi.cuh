// nm_multi.cuh -- channel-sharded multi-GPU upload.
//
// Every rank owns a contiguous block of channels of the recording.  After re-referencing all hot-path features are
// per-channel, so the only exchange on the data path is the common-average reference.  The upload is asynchronous and sliced
// in time like the single-GPU one:
//   nm_upload_begin_f32       H2D of the local shard in time slices (copy stream)
//   nm_upload_slice_sums(k)   local per-sample group sums S_g[t] of slice k on the reduction stream (after the slice has landed)
//   (host, per slice k)       all-reduce(sum) of S[:, slice k] over the ranks, enqueued on the reduction stream (the host hands
//                             the stream to its collective library: torch.cuda.ExternalStream + NCCL), then
//   nm_upload_slice_reduced   records "slice k reduced" on that stream
//   nm_upload_finish          switches the pipeline to the sharded re-reference: nm_run_windows re-references slice k on the
//                             compute stream right before the first chunk that needs it, after waiting for that event
// so transfers, reductions and window kernels of different slices overlap and no call blocks the host.  The result blocks
// are collected by the host (shared page-locked matrix or NCCL gather, parallel.py).  Included at the end of nm_pipeline.cu.
#pragma once

extern "C" int nm_upload_begin_f32(nm_pipeline* p, const float* data, long long n_samples, long long pitch) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized, "call nm_finalize first");
    NM_CHECK(data && n_samples >= p->Win && pitch >= n_samples, "bad recording geometry");
    NM_CHECK(p->G > 0, "nm_upload_begin_f32 needs a re-reference with at least one channel group");
    cudaSetDevice(p->device);
    nm_upload_join(p);  // (a deferred single-GPU upload may still be staging)
    nm_stream_release(p);
    const int n_slices = n_samples >= NM_UPLOAD_MIN_PIPELINED ? NM_UPLOAD_SLICES : 1;
    if (nm_stage_slices(p, data, false, n_samples, pitch, n_slices)) return -1;
    p->gsum_pitch = (n_samples + 1) & ~1LL;
    if (p->d_gsum.ensure((size_t)p->G * p->gsum_pitch * sizeof(double) + 16)) return -1;
    // the reduction stream must not run ahead of the previous run's readers of d_gsum (compute stream)
    NM_CUDA_CHECK(cudaStreamWaitEvent(p->red_stream, p->ev_sync, 0));
    NM_CUDA_CHECK(cudaGetLastError());
    p->upload_pending = true;
    p->have_data = false;
    return 0;
}

// local group sums of slice k on the reduction stream (enqueued per slice, right before the host's all-reduce of that slice, so that
// the stream order is  sums(0), reduce(0), sums(1), reduce(1), ...  and slice 0 does not wait for the last transfer)
extern "C" int nm_upload_slice_sums(nm_pipeline* p, int k) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending && k >= 0 && k < p->n_slices, "no such slice");
    cudaSetDevice(p->device);
    NmPrepArgs a = nm_prep_args(p);
    a.gsum_pitch = p->gsum_pitch;
    a.t0 = (long long)k * p->slice_len;
    a.t1 = std::min<long long>(p->T, a.t0 + p->slice_len);
    const unsigned grid = (unsigned)((a.t1 - a.t0 + 31) / 32);
    NM_CUDA_CHECK(cudaStreamWaitEvent(p->red_stream, p->slice_ev[k], 0));
    NM_LAUNCH(nm_gsum_kernel, dim3(grid), dim3(NM_PREP_THREADS), nm_prep_smem_bytes(), p->red_stream, a, p->d_gsum.as<double>());
    p->launches++;
    NM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int nm_group_sums_device_ptr(nm_pipeline* p, void** ptr, long long* n_values) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending, "no sharded upload in progress");
    if (ptr) *ptr = p->d_gsum.p;
    if (n_values) *n_values = (long long)p->G * p->gsum_pitch;
    return 0;
}

extern "C" int nm_upload_slices(nm_pipeline* p, int* n_slices, long long* slice_len, int* n_groups, long long* group_pitch) {
    NM_P_CHECK(p);
    if (n_slices) *n_slices = p->n_slices;
    if (slice_len) *slice_len = p->slice_len;
    if (n_groups) *n_groups = p->G;
    if (group_pitch) *group_pitch = p->gsum_pitch;
    return 0;
}

extern "C" int nm_side_stream_handle(nm_pipeline* p, void** cuda_stream) {
    NM_P_CHECK(p);
    if (cuda_stream) *cuda_stream = (void*)p->red_stream;
    return 0;
}

extern "C" int nm_upload_slice_reduced(nm_pipeline* p, int k) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending && k == p->slices_reduced && k < p->n_slices, "slices must be reported in order (got %d, expected %d)", k,
             p->slices_reduced);
    cudaSetDevice(p->device);
    NM_CUDA_CHECK(cudaEventRecord(p->red_ev[k], p->red_stream));
    p->slices_reduced = k + 1;
    return 0;
}

extern "C" int nm_upload_finish(nm_pipeline* p) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending, "no sharded upload in progress");
    NM_CHECK(p->slices_reduced == p->n_slices, "%d of %d slices reduced", p->slices_reduced, p->n_slices);
    p->upload_pending = false;
    p->have_data = true;
    p->resident_uses_gsum = true;
    return 0;
}
