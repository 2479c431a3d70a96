/* nmb200.h -- C ABI of libnmb200.so: the B200 (sm_100a) implementation of py_neuromodulation's
 * per-window feature-extraction hot path.
 *
 * The reference (py_neuromodulation v0.1.4) is pure Python and has NO foreign-function
 * interface; this ABI is new.  Each entry point cites the reference interface whose work it
 * takes over (paths relative to the reference's py_neuromodulation/ directory).  A reference
 * maintainer binds it with ctypes -- see INTEGRATION.md for the stub.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and -1 on
 * failure, in which case nm_last_error() describes the problem (thread local).  The caller
 * owns all host memory; device memory is owned by the opaque pipeline handle.
 * All feature values are float64, laid out row-major as (n_windows, n_features); the column
 * of every value is chosen by the caller through the `colmap` arrays (-1 = do not emit), so
 * the host side keeps the reference's dict-insertion column order without a permutation pass.
 *
 * A pipeline is built for ONE window length W (samples) and one channel set:
 *   create -> set_* (preprocessing) -> add_* (feature families) -> finalize
 *   -> upload -> run_windows (offline batch, stream/stream.py:280-330)
 *   or process_window (one window at a time, stream/data_processor.py:238-311).
 */
#ifndef NMB200_H
#define NMB200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NMB200_ABI_VERSION 1

typedef struct nm_pipeline nm_pipeline;

/* ---- library ------------------------------------------------------------------------------ */
const char* nm_last_error(void);
int nm_abi_version(void);
int nm_device_count(int* count);
/* pinned host memory for recordings / result matrices (host<->device copies at full PCIe rate) */
int nm_host_alloc(void** ptr, long long bytes);
int nm_host_free(void* ptr);

/* ---- pipeline life cycle ------------------------------------------------------------------ */
/* n_raw_rows: rows of the array handed to Stream.run / DataProcessor.process (all channels);
 * n_ch: rows used for features (stream/data_processor.py:141-160 `feature_idx`);
 * window_samples: W; n_features: F, number of output columns. */
int nm_pipeline_create(int device, int n_raw_rows, int n_ch, int window_samples, int n_features, nm_pipeline** out);
void nm_pipeline_destroy(nm_pipeline* p);
int nm_finalize(nm_pipeline* p);
/* forget cross-window state (burst history, normaliser history): what constructing a new
 * DataProcessor does in stream/stream.py:233-242 */
int nm_reset_state(nm_pipeline* p);

/* ---- preprocessing (stream/data_processor.py:253-257, processing/data_preprocessor.py:74-84) */
/* pick[n_ch]: raw row of each feature channel (np.nan_to_num(data)[feature_idx, :]) */
int nm_set_pick(nm_pipeline* p, const int* pick);
/* ReReferencer.process == ref_matrix @ data (processing/rereference.py:88-102), handed over
 * factored: y_i = sum_g gcoef[i,g] * S_g + sum_k sp_val[k] * x[sp_col[k]],  S_g = sum of the
 * channels with group_of[j] == g.  n_groups <= 8. */
int nm_set_reref(nm_pipeline* p, int n_groups, const int* group_of, const double* gcoef,
                 const int* sp_ptr, const int* sp_col, const double* sp_val);
/* NotchFilter.process: zero-phase FIR with reflect-limited padding (filter/notch_filter.py:78-93) */
int nm_set_notch(nm_pipeline* p, const double* taps, int n_taps);
/* Arithmetic of the linear FIR families (notch, band-pass power): 0 = float64 (default: agrees with the float64 reference to
 * ~1e-12), 1 = float32 inside the FFT convolution (faster; north-star tolerance 1e-5 relative; moments and outputs stay
 * float64), 2 = the same arithmetic on PACKED float32 pairs (two channel pairs per item, Blackwell add/mul/fma.f32x2; measured
 * slower than 1 on B200, kept for comparison).  Pipelines with threshold / peak decisions downstream of the notch (bursts, sharp
 * waves, raw normaliser) keep the notch in float64 regardless. */
int nm_set_precision(nm_pipeline* p, int float32_linear);
/* Kernel organisation of the window chain.  0 (default): one kernel per stage -- re-reference (once per recording), notch (+ Hjorth /
 * line length / raw), segment DFTs, band-pass bank -- the notched rows travel through HBM / L2.  1: ONE persistent kernel per
 * (window, channel pair) (csrc/nm_fused.cuh): raw rows staged by cp.async.bulk + mbarrier, re-reference folded into the load, notch
 * -> scan -> DFT band features -> band-pass bank without the re-referenced recording or the notched rows ever existing in HBM.
 * Same results (tests run both); measured 15 % slower on B200 at float64 because the merged kernel runs every phase at the bank's
 * occupancy, hence opt-in.  2: the "front" kernel -- the same bulk-copy staged load with the folded
 * re-reference, notch + Hjorth / line length / raw (+ the segment DFTs with NMB200_FRONT_DFT=1) in one kernel, band-pass bank
 * separate; measured 2.5 % slower than 0 (the fold repeats per window what the re-reference kernel does once per sample).
 * -1: take the choice from the environment (NMB200_FUSED, default 0).  Independently of this choice the staged notch kernel of
 * mode 0 receives its float64 rows through the TMA engine (cp.async.bulk + mbarrier; NMB200_NOTCH_TMA=0 restores the register
 * prefetch). */
int nm_set_fused(nm_pipeline* p, int mode);

/* RawNormalizer (processing/normalization.py:30-111, type "raw"): window 0 passes through and seeds the per-channel
 * history; window g >= 1 appends its last add_samples = int(sfreq / rate) preprocessed samples, is normalised against the
 * whole history (method 0 mean, 1 median, 2 zscore, 3 zscore-median, 4 minmax = scikit-learn MinMaxScaler, 5 robust = RobustScaler;
 * the order statistics come from the sliding quantile kernel of the burst thresholds), clipped (clip == 0: off) and the history is
 * trimmed to n_keep - 1 samples
 * (n_keep = int(normalization_time_s * sfreq)).  Runs after the notch / re-reference, before every feature. */
int nm_set_raw_normalizer(nm_pipeline* p, int method, double clip, int n_keep, int add_samples);

/* PreprocessingFilter (processing/filter_preprocessing.py:44-94): appends one single 'same' FIR stage (odd-length,
 * symmetric taps; stages may differ in length); the stages are applied in the order added to every window before
 * the notch. */
int nm_add_prefilter(nm_pipeline* p, const double* taps, int n_taps);

/* Resampler (processing/resample.py:28-60, mne.filter.resample(x, up = resample_freq_hz / sfreq, down = 1)): FFT resampling of
 * a window row is a linear map; `op` is that map as a dense row-major (window_samples x n_in) float64 matrix designed by the
 * host (processing/resample.py::resample_operator here).  Windows are then cut from the recording, pre-filtered and notched at
 * n_in samples and every row is mapped to the pipeline's `window_samples` before re-referenced rows reach the raw normaliser and
 * the feature families (the reference's fixed preprocessor order).  Call before nm_add_prefilter / nm_set_notch.  Everything
 * downstream keeps the ORIGINAL sampling rate, like the reference (stream/data_processor.py:55,77-81). */
int nm_set_resampler(nm_pipeline* p, int n_in, const double* op, int fft_decim);
/* fft_decim: 0, or the integer factor D >= 2 when `op` is MNE's default FFT down-sampler by D (npad="auto", reflect-limited pad,
 * boxcar window).  The library then checks entries of `op` against the closed form and, if the padded length is one of the
 * specialised transform sizes, runs the resampler as an FFT convolution with an ideal low-pass whose epilogue stores every D-th
 * sample (two transforms per channel pair) instead of the dense GEMM.  0 always selects the GEMM. */

/* NaN re-insertion (stream/data_processor.py:297-306): columns [col_ptr[r], col_ptr[r+1]) of `cols`
 * become NaN in every window where raw row r contains a NaN */
int nm_set_nan_columns(nm_pipeline* p, const int* col_ptr, const int* cols);

/* ---- feature families ---------------------------------------------------------------------- */
/* Hjorth / Raw / LineLength (features/hjorth_raw.py:24-57, features/linelength.py:11-21);
 * colmap[n_ch*5]: activity, mobility, complexity, raw, linelength */
int nm_add_scan(nm_pipeline* p, int hjorth, int raw, int linelength, const int* colmap);

/* FFT / Welch / STFT band features (features/oscillatory.py:90-119,150-182,215-250) described as
 * nseg segment DFTs of nper samples; see csrc/nm_spec.cuh for the three parameterisations. */
typedef struct {
    int nper, nseg, hop, start;
    int ext_even, ext_len;
    int detrend;
    int power;          /* 0: |Z|*scale   1: |Z|^2*scale, interior bins doubled */
    double scale;
    int log;
    int keep_segments;  /* 1: estimators over the (bin, segment) matrix (STFT) */
    int n_bands;
    int est_mask;       /* bit0 mean, bit1 median, bit2 std, bit3 max */
    int want_spectrum;
    const double* win;  /* [nper] or NULL */
    const int* band_lo; /* [n_bands] first bin  */
    const int* band_hi; /* [n_bands] one past the last bin */
    const int* colmap;  /* [n_ch * (n_bands*4 + nper/2+1)]: band b estimator e at b*4+e, then spectrum bins */
} nm_spectral_cfg;
int nm_add_spectral(nm_pipeline* p, const nm_spectral_cfg* cfg);

/* BandPower: MNEFilter.filter_data + tail variance features (filter/mne_filter.py:82-128,
 * features/bandpower.py:165-207); taps[n_bands*n_taps] symmetric; colmap[n_ch*n_bands*3] */
int nm_add_bandpower(nm_pipeline* p, int n_bands, const double* taps, int n_taps, const int* seglen,
                     int activity, int mobility, int complexity, int log_transform, const int* colmap);

/* Bursts (features/bursts.py:149-298): band-pass bank -> analytic-signal envelope -> quantile of the
 * last ring_samples envelope samples -> run-length features.
 * colmap[n_ch*n_bands*6]: duration_mean, duration_max, amplitude_mean, amplitude_max, burst_rate_per_s, in_burst.
 * qlo/qgamma give numpy's 'linear' quantile position for every possible history length:
 * the host passes a callback-free closed form: quantile q in [0,1]; the library evaluates
 * numpy's virtual index n*q + (1 - q) - 1 in float64 exactly as numpy does. */
int nm_add_bursts(nm_pipeline* p, int n_bands, const double* taps, int n_taps, int samples_overlap,
                  int ring_samples, double quantile, double sfreq, double segment_length_s, const int* colmap);

/* SharpwaveAnalyzer (features/sharpwaves.py:225-465).  feat_ids/est_ids[n_combo] list the
 * (feature, estimator) pairs in output order; colmap[n_ch*n_filters*(n_combo+1)*2]: slot (f*(n_combo+1)+k)*2+pol,
 * k == n_combo being num_peaks; pol 0 holds the joined value when pair_estimator is set, else pol 0/1 = Peak/Trough pass.  Feature ids: 0 peak_left 1 peak_right 2 num_peaks 3 trough 4 width 5 prominence
 * 6 interval 7 decay_time 8 rise_time 9 sharpness 10 rise_steepness 11 decay_steepness 12 slope_ratio.
 * Estimator ids: 0 mean 1 median 2 max 3 min 4 var. */
int nm_add_sharpwave(nm_pipeline* p, int n_filters, const double* taps, int n_taps, int dist_peaks, int dist_troughs,
                     int sharp_offset, double ms_per_sample, int n_combo, const int* feat_ids, const int* est_ids,
                     int pair_estimator, int want_num_peaks, const int* colmap);

/* FeatureNormalizer (processing/normalization.py:81-111): rolling normalisation of the columns listed in
 * `cols` over the previous n_keep windows (current included); method 0 mean, 1 median, 2 zscore,
 * 3 zscore-median, and the scikit-learn transformers the reference wraps (processing/normalization.py:58-70,173-190),
 * restated: 4 minmax (MinMaxScaler), 5 robust (RobustScaler), 6 quantile (QuantileTransformer(n_quantiles=300), n_keep <= 300);
 * clip <= 0 disables clipping. */
int nm_add_feature_normalizer(nm_pipeline* p, int method, double clip, int n_keep, int n_cols, const int* cols);

/* ---- data path ------------------------------------------------------------------------------ */
/* Recording (n_raw_rows, n_samples), row pitch in elements; copies host -> device and runs the
 * window-independent preprocessing (nan_to_num, pick, re-reference).  Recordings of >= 65 536 samples are copied
 * ASYNCHRONOUSLY in time slices on a second stream (each slice is re-referenced right before the first chunk of
 * windows that needs it, so the transfer overlaps the window kernels): `data` must stay valid and unchanged until
 * the next nm_run_windows / nm_synchronize on this pipeline has returned.  Page-locked buffers are read by the copy engine
 * directly; PAGEABLE buffers are packed slice by slice into two page-locked staging buffers by a host thread of the library,
 * so the call returns at once and the staging overlaps the window kernels too (NMB200_DEFERRED_UPLOAD=0: the driver's
 * synchronous staged copy instead). */
int nm_upload_f32(nm_pipeline* p, const float* data, long long n_samples, long long pitch);
int nm_upload_f64(nm_pipeline* p, const double* data, long long n_samples, long long pitch);
/* RawDataGenerator windows (stream/generator.py:41-53): window k = samples [starts[k], starts[k]+W).
 * out_host may be NULL (results stay on the device; use nm_download).  With out_host and no feature normaliser the
 * rows of every finished chunk are copied back while the next chunk computes; the call returns when all rows are
 * on the host. */
int nm_run_windows(nm_pipeline* p, const long long* starts, int n_windows, double* out_host);
int nm_download(nm_pipeline* p, double* out_host, int n_windows);
/* Row pitch (in elements, >= n_features; 0 = dense) of the HOST matrix that nm_run_windows / nm_download write: lets every rank
 * of a channel-sharded run copy its (n_windows x F_local) block straight into its column range of one shared host matrix. */
int nm_set_output_pitch(nm_pipeline* p, long long pitch_elems);
/* page-lock / unlock caller-owned host memory (e.g. a POSIX shared-memory segment mapped by all ranks of a node) */
int nm_host_register(void* ptr, long long bytes);
int nm_host_unregister(void* ptr);
/* DataProcessor.process for one (n_raw_rows, W) float64 window; keeps cross-window state. */
int nm_process_window(nm_pipeline* p, const double* window, double* out_features);

/* DataPreprocessor.process_data for one window (processing/data_preprocessor.py:74-84): nan_to_num -> pick ->
 * re-reference -> notch; out_rows (n_ch, W) float64.  Feeds user-defined Python features that run next to the
 * GPU families (features/feature_processor.py:52-53) and the stand-alone ReReferencer class. */
int nm_preprocess_window(nm_pipeline* p, const double* window, double* out_rows);

/* ---- stand-alone FIR application -------------------------------------------------------------- */
/* MNEFilter.filter_data (mode 0: scipy fftconvolve 'same', filter/mne_filter.py:82-128) and NotchFilter.process
 * (mode 1: reflect-limited zero-phase overlap-add, filter/notch_filter.py:78-93) for callers that use those
 * classes outside a pipeline.  data (n_ch, n_samples) -> out (n_ch, n_filters, n_samples), float64. */
int nm_fir_apply(int device, const double* taps, int n_filters, int n_taps, int mode, const double* data, int n_ch,
                 int n_samples, double* out);

/* ---- measurement ---------------------------------------------------------------------------- */
/* CUDA events on the pipeline's stream */
int nm_timer_start(nm_pipeline* p);
int nm_timer_stop(nm_pipeline* p, double* elapsed_ms);
/* re-run nan_to_num / pick / re-reference on the recording already resident on the device (benchmarks) */
int nm_prepare_resident(nm_pipeline* p);
int nm_synchronize(nm_pipeline* p);
/* per-family kernel timing: when enabled every launch is bracketed by CUDA events on the pipeline's stream and
 * host-synchronised (profiling runs only).  nm_get_profile fills ms[i] / launches[i] for family i and returns the
 * number of families: 0 prep, 1 notch, 2 scan, 3 spectral, 4 bandpower, 5 sharpwave, 6 burst envelope,
 * 7 burst threshold, 8 burst features, 9 normaliser, 10 nan. */
int nm_set_profiling(nm_pipeline* p, int enabled);
int nm_get_profile(nm_pipeline* p, double* ms, long long* launches, int n);
/* windows per kernel launch (chunk size chosen so that the notched chunk stays L2 resident) */
int nm_chunk_windows(nm_pipeline* p);
/* burst thresholds: 1 (default) = incremental sliding order statistics, 0 = re-select from the whole history for every
 * window (slower reference implementation of the same kernel; results are identical).  Resets the burst state. */
int nm_set_burst_threshold_mode(nm_pipeline* p, int incremental);
/* statistics of the incremental thresholds since the last reset, summed over (channel, band) rows: bracket rebuilds
 * and windows that fell back to the direct selection */
int nm_burst_threshold_stats(nm_pipeline* p, long long* rebuilds, long long* direct_windows);
/* text description of the launch plan (one line per family: kernel, transform size, threads, shared memory);
 * writes at most n-1 characters + NUL into buf and returns the full length */
int nm_describe_plan(nm_pipeline* p, char* buf, int n);
/* number of kernels this library launched since the pipeline was created */
long long nm_kernel_launches(nm_pipeline* p);
/* device pointer / geometry of the last result matrix (for NCCL gathers by the host side) */
int nm_result_device_ptr(nm_pipeline* p, void** ptr, long long* n_rows, int* n_cols);
int nm_stream_handle(nm_pipeline* p, void** cuda_stream);
/* Channel-sharded multi-GPU runs (asynchronous, sliced in time like nm_upload_f32).  nm_upload_begin_f32 enqueues the H2D
 * slices of the local shard.  Per slice k (in order): nm_upload_slice_sums enqueues the local per-sample group sums S
 * (n_groups x group_pitch float64 at nm_group_sums_device_ptr) on the stream returned by nm_side_stream_handle, the host
 * all-reduces S[:, k*slice_len : ...] across ranks ON THAT STREAM (slice geometry: nm_upload_slices) and reports
 * nm_upload_slice_reduced;
 * and closes with nm_upload_finish; nm_run_windows then re-references each slice with the global sums right before the first
 * chunk of windows that needs it.  `data` must stay valid until the next nm_run_windows / nm_synchronize has returned. */
int nm_upload_begin_f32(nm_pipeline* p, const float* data, long long n_samples, long long pitch);
int nm_group_sums_device_ptr(nm_pipeline* p, void** ptr, long long* n_values);
int nm_upload_slices(nm_pipeline* p, int* n_slices, long long* slice_len, int* n_groups, long long* group_pitch);
int nm_side_stream_handle(nm_pipeline* p, void** cuda_stream);
int nm_upload_slice_sums(nm_pipeline* p, int slice);
int nm_upload_slice_reduced(nm_pipeline* p, int slice);
int nm_upload_finish(nm_pipeline* p);

/* ---- streaming entry (csrc/nm_stream.cuh): one window at a time, the way the reference is driven by a live source
 * (stream/stream.py:280-330 calls DataProcessor.process per batch; stream/mnelsl_stream.py feeds it).  A ring of page-locked slots:
 * the producer writes the next (n_raw_rows x window) block into a slot's input, nm_stream_submit enqueues H2D + every kernel + D2H
 * of the feature row WITHOUT blocking, nm_stream_wait blocks on that slot only.  The launch sequence of a window is captured into
 * a CUDA graph per slot and replayed (window counters of the stateful stages are patched into the graph); stateful stages (bursts
 * history, raw / feature normaliser) advance exactly as in nm_process_window, which is the synchronous wrapper of these calls.
 * input_f32: slot samples are float32 instead of float64.  use_graph: 1 graph replay, 0 eager launches, -1 environment
 * (NMB200_STREAM_GRAPH, default 1).  Batched entry points (nm_upload_*, nm_run_windows) close the stream. */
int nm_stream_open(nm_pipeline* p, int n_slots, int input_f32, int use_graph);
int nm_stream_input(nm_pipeline* p, int slot, void** ptr, long long* bytes);
int nm_stream_submit(nm_pipeline* p, int slot);
int nm_stream_wait(nm_pipeline* p, int slot, const double** features);
int nm_stream_stats(nm_pipeline* p, long long* windows, long long* graph_launches, long long* graph_nodes, long long* patched);
int nm_stream_close(nm_pipeline* p);

/* ---- collectives inside the library (csrc/nm_comm.cuh): NCCL over NVLink / NVSwitch, one process per GPU, no torch.
 * The reference has no multi-GPU path; these entry points are what a binding adds for SURVEY.md section 8e (channel shards,
 * one all-reduce of the common-average sums, one gather of the result blocks).  NCCL is dlopen'ed ("libnccl.so.2"). */
typedef struct nm_comm nm_comm;
/* rank 0 creates the 128-byte NCCL unique id; the application hands it to every rank (MPI, file, TCP store, ...) */
int nm_comm_unique_id(unsigned char* id128);
int nm_comm_create(const unsigned char* id128, int rank, int world, int device, nm_comm** out);
void nm_comm_destroy(nm_comm* c);
int nm_comm_rank(const nm_comm* c);
int nm_comm_size(const nm_comm* c);
long long nm_comm_collectives(const nm_comm* c);        /* collective launches issued so far (grouped calls count once) */
int nm_comm_barrier(nm_comm* c);
int nm_comm_allreduce_max(nm_comm* c, double* value);   /* in place, host value: "time on the device, max over ranks" */
/* nm_upload_begin_f32 + per slice {local group sums, ncclAllReduce(sum) of that slice on the reduction stream} + nm_upload_finish:
 * returns without blocking the host; nm_run_windows waits per slice on an event.  The pipeline's re-reference must have been set
 * from the GLOBAL channel table (nm_set_reref with group coefficients -1/(n_global - 1)). */
int nm_upload_sharded_f32(nm_pipeline* p, nm_comm* c, const float* data, long long n_samples, long long pitch);
/* result blocks of all ranks -> rank 0: out_host (n_windows x sum(widths)), rank-major column blocks; widths[r] = feature columns
 * of rank r (uneven shards allowed).  Other ranks pass out_host = NULL.  Blocking on rank 0 (returns after the D2H). */
int nm_gather_results(nm_pipeline* p, nm_comm* c, int n_windows, const int* widths, double* out_host);
/* sharded counterpart of nm_prepare_resident: local group sums of the resident shard, ncclAllReduce, window-independent preprocessing */
int nm_prepare_resident_sharded(nm_pipeline* p, nm_comm* c);

#ifdef __cplusplus
}
#endif
#endif /* NMB200_H */
