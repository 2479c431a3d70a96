// nm_multi.cuh -- channel-sharded multi-GPU upload.
//
// Every rank owns a contiguous block of channels of the recording.  After re-referencing all hot-path
// features are per-channel, so the only exchange on the data path is the common-average reference:
//   nm_upload_begin_f32   H2D of the local shard + local per-sample group sums  S_g[t]
//   (host)                one all-reduce(sum) of S over the ranks (G x T float64; NCCL over NVLink)
//   nm_upload_finish      re-reference with the global sums (coefficients were computed by the host
//                         from the GLOBAL channel table)
// and one gather of the (n_windows x F_local) result blocks at the end (done by the host through
// nm_result_device_ptr).  Included at the end of nm_pipeline.cu.
#pragma once

extern "C" int nm_upload_begin_f32(nm_pipeline* p, const float* data, long long n_samples, long long pitch) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized, "call nm_finalize first");
    NM_CHECK(data && n_samples >= p->W && pitch >= n_samples, "bad recording geometry");
    NM_CHECK(p->G > 0, "nm_upload_begin_f32 needs a re-reference with at least one channel group");
    cudaSetDevice(p->device);
    if (nm_stage_raw(p, data, false, n_samples, pitch)) return -1;
    p->gsum_pitch = (n_samples + 1) & ~1LL;
    if (p->d_gsum.ensure((size_t)p->G * p->gsum_pitch * sizeof(double))) return -1;
    NmPrepArgs a = nm_prep_args(p);
    a.gsum_pitch = p->gsum_pitch;
    const int threads = NM_ROW_THREADS;
    const unsigned grid = (unsigned)((n_samples + threads - 1) / threads);
    NM_LAUNCH(nm_gsum_kernel, dim3(grid), dim3(threads), 0, p->stream, a, p->d_gsum.as<double>());
    p->launches++;
    NM_CUDA_CHECK(cudaGetLastError());
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->upload_pending = true;
    return 0;
}

extern "C" int nm_group_sums_device_ptr(nm_pipeline* p, void** ptr, long long* n_values) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending, "no sharded upload in progress");
    if (ptr) *ptr = p->d_gsum.p;
    if (n_values) *n_values = (long long)p->G * p->gsum_pitch;
    return 0;
}

extern "C" int nm_upload_finish(nm_pipeline* p) {
    NM_P_CHECK(p);
    NM_CHECK(p->upload_pending, "no sharded upload in progress");
    cudaSetDevice(p->device);
    NmPrepArgs a = nm_prep_args(p);
    a.gsum_ext = p->d_gsum.as<double>();
    a.gsum_pitch = p->gsum_pitch;
    const int threads = NM_ROW_THREADS;
    const unsigned grid = (unsigned)((p->T + threads - 1) / threads);
    NM_LAUNCH(nm_prep_kernel, dim3(grid), dim3(threads), 0, p->stream, a);
    p->launches++;
    NM_CUDA_CHECK(cudaGetLastError());
    p->upload_pending = false;
    p->have_data = true;
    p->resident_uses_gsum = true;
    return 0;
}
