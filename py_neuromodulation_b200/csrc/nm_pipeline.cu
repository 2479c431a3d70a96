// nm_pipeline.cu -- the C ABI (include/nmb200.h): pipeline object, plan construction, launches.
//
// Data flow of one offline run (Stream.run fast path):
//
//   host recording (C_all x T, f32/f64) --H2D--> raw
//   raw --nm_prep_kernel--> xr (C x T, f64): nan_to_num, pick, re-reference       [once]
//   for each chunk of windows (sized so that the notched chunk stays L2-resident):
//       xr windows --nm_fir_kernel<store>--> Y chunk (notch, reflect-limited)       [if notch]
//       Y / xr rows --scan / spectral / band-power / bursts / sharp-wave kernels--> out columns
//   out --normaliser--> out, NaN re-insertion                                      [whole run]
//   out (n_windows x F, f64) --D2H--> host
#include <condition_variable>
#include <cstdarg>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/nmb200.h"
#include "nm_host.h"
#include "nm_prep.cuh"
#include "nm_firbank.h"
#include "nm_scan.cuh"
#include "nm_specx.cuh"
#include "nm_fused.cuh"
#include "nm_bursts.cuh"
#include "nm_sharpwave.cuh"
#include "nm_norm.cuh"
#include "nm_rawnorm.cuh"
#include "nm_resample.cuh"
#include "nm_stream.h"

// ------------------------------------------------------------------------------- errors
static thread_local char g_err[1024] = "";
void nm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* nm_last_error(void) { return g_err; }
extern "C" int nm_abi_version(void) { return NMB200_ABI_VERSION; }
extern "C" int nm_device_count(int* count) {
    NM_CHECK(count, "count is NULL");
    NM_CUDA_CHECK(cudaGetDeviceCount(count));
    return 0;
}
extern "C" int nm_host_alloc(void** ptr, long long bytes) {
    NM_CHECK(ptr && bytes >= 0, "bad arguments");
    NM_CUDA_CHECK(cudaMallocHost(ptr, (size_t)(bytes ? bytes : 16)));
    return 0;
}
extern "C" int nm_host_free(void* ptr) {
    NM_CUDA_CHECK(cudaFreeHost(ptr));
    return 0;
}

// ------------------------------------------------------------------------------- families
struct SpectralFam {
    nm_spectral_cfg cfg;
    FftPlanHost fft;
    DevBuf d_win, d_lo, d_hi, d_colmap;
    int k0 = 0, nk = 0, per_ch = 0;
    bool fast = false;  // register-blocked three-pass plan (nm_specx.cuh) instead of the generic mixed-radix kernel
    int nsegv() const { return cfg.keep_segments ? cfg.nseg : 1; }
    int nbuf() const { return cfg.nper == 1000 ? NmSx1000::NBUF : (cfg.nper == 2000 ? NmSx2000::NBUF : NmSx500::NBUF); }
    int threads() const { return !fast ? NM_FFT_THREADS : (cfg.nper == 500 ? NmSx500::NT : NmSx1000::NT); }
    size_t smem() const { return fast ? nm_specx_smem_bytes(nbuf(), nk, nsegv()) : nm_spec_smem_bytes(cfg.nper, fft.generic, nk, nsegv()); }
    void (*kernel() const)(NmSpecArgs) {
        if (!fast) return nm_spec_kernel;
        if (cfg.nper == 1000) return nm_specx_kernel<NmSx1000>;
        if (cfg.nper == 2000) return nm_specx_kernel<NmSx2000>;
        return nm_specx_kernel<NmSx500>;
    }
};

struct BandpowerFam {
    FirBank bank;
    DevBuf d_seglen, d_colmap;
    std::vector<int> h_seglen;
    int act = 0, mob = 0, comp = 0, logt = 0;
    // the register epilogue (activity only) keeps its reduction scratch inside the free `work` buffer
    size_t epi_smem() const { return NmEpiBandpower::smem_bytes(NM_FFT_THREADS); }
};

enum { NM_PROF_PREP = 0, NM_PROF_NOTCH, NM_PROF_SCAN, NM_PROF_SPEC, NM_PROF_BANDPOWER, NM_PROF_SHARPWAVE, NM_PROF_BURST_ENV,
       NM_PROF_BURST_THR, NM_PROF_BURST_FEAT, NM_PROF_NORM, NM_PROF_NAN, NM_PROF_FUSED, NM_PROF_RESAMPLE, NM_PROF_N };

struct nm_pipeline {
    int device = 0, C_all = 0, C = 0, W = 0, F = 0;
    // raw_resampling (nm_resample.cuh): windows are cut, pre-filtered and notched at Win samples, every row is then mapped to W
    // samples by the host-designed operator; without a resampler Win == W
    int Win = 0, Wpi = 0;
    bool has_resampler = false;
    int rs_pitch = 0;                 // row pitch of the transposed operator
    DevBuf d_rt, d_rs, d_roff;        // operator (Win x rs_pitch), resampled chunk rows (chunk x C x Wp), their window offsets
    std::unique_ptr<FirBank> rs_bank; // integer down-sampling: ideal low-pass on the FFT-convolution kernel + decimating store
    int rs_decim = 0;
    bool finalized = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // copy engine overlap: H2D of the recording in time slices and D2H of finished chunks run on `copy_stream` while the
    // window kernels run on `stream`; slices are re-referenced lazily, right before the first chunk that needs them
    cudaStream_t copy_stream = nullptr;
    // deferred upload of a PAGEABLE recording: a host thread packs time slices into two page-locked staging
    // buffers and enqueues their H2D copies while the caller already launches window kernels; `up_enqueued` slices have their
    // events recorded (a cudaStreamWaitEvent on an event that has not been recorded yet would not wait)
    std::thread up_thread;
    std::mutex up_mu;
    std::condition_variable up_cv;
    int up_enqueued = 0;
    bool up_active = false, up_failed = false;
    // pageable result matrix: finished chunks are copied into this page-locked buffer asynchronously and moved on to the caller's
    // rows by the host while later chunks still compute (a D2H straight into pageable memory blocks the launching thread per chunk)
    void* out_stage = nullptr;
    size_t out_stage_bytes = 0;
    std::vector<cudaEvent_t> d2h_ev;
    void* up_stage[2] = {nullptr, nullptr};
    size_t up_stage_bytes = 0;
    cudaEvent_t up_free[2] = {nullptr, nullptr};
    cudaEvent_t ev_sync = nullptr;
    // independent feature families of a chunk (spectral + band power | sharp waves | bursts) run on side streams between a
    // fork after the notch and a join before the next chunk: their kernels are latency / occupancy limited in different ways
    cudaStream_t side[3] = {nullptr, nullptr, nullptr};
    cudaStream_t red_stream = nullptr;  // channel-sharded uploads: per-slice group sums + the host's all-reduce (nm_multi.cuh)
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> slice_ev, chunk_ev, red_ev;  // red_ev[k]: group sums of slice k are all-reduced (sharded runs)
    long long slice_len = 0;
    int n_slices = 0, slices_prepped = 0, slices_reduced = 0;
    int n_sm = 1, smem_max = 48 * 1024;
    long long launches = 0;
    // optional per-family kernel timing (nm_set_profiling): events around every launch, host-synchronised
    bool profiling = false;
    double prof_ms[NM_PROF_N] = {0};
    long long prof_cnt[NM_PROF_N] = {0};
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    void prof_begin() {
        if (profiling) cudaEventRecord(pe0, stream);
    }
    void prof_end(int fam) {
        if (!profiling) return;
        cudaEventRecord(pe1, stream);
        cudaEventSynchronize(pe1);
        float ms = 0;
        cudaEventElapsedTime(&ms, pe0, pe1);
        prof_ms[fam] += ms;
        prof_cnt[fam] += 1;
    }

    // preprocessing
    std::vector<int> pick;
    DevBuf d_pick, d_group_of, d_gcoef, d_sp_ptr, d_sp_col, d_sp_val;
    int G = 0;
    bool has_reref = false;
    // fused window kernel (nm_fused.cuh): the re-reference must fold into the load as  x_i = d_i * raw_i + g_i * S  (at most one
    // group sum, no off-diagonal remainder); decided at nm_finalize, switchable with NMB200_FUSED=0 for A/B measurements
    bool reref_foldable = true, fused = false, force_xr = false, fused_bp = false;
    bool notch_tma = false;  // staged path: the notch kernel gets its rows through cp.async.bulk + mbarrier (nm_notchx_kernel)
    bool front = false;   // `fused` through nm_front_kernel (notch + scan + segment DFT; the bank stays a kernel of its own)
    int fused_mode = -1;  // nm_set_fused: 0 staged kernels, 1 whole chain in one kernel, 2 front kernel, -1 environment (NMB200_FUSED, default 0)
    std::vector<double> h_dfold, h_gfold;
    DevBuf d_dfold, d_gfold;
    int fused_sx = 0;                       // segment length of the in-kernel DFT families (0: none)
    std::vector<int> fused_spec;            // indices into `spectral` of the families computed inside the fused kernel
    std::vector<NmSpecArgs> h_fused_spec;
    DevBuf d_fused_spec;                    // their NmSpecArgs blocks (out.row0 == 0; the launch passes the batch's first row)
    std::unique_ptr<FirBank> notch;
    std::vector<double> notch_taps;
    // PreprocessingFilter (processing/filter_preprocessing.py): single-filter 'same' FIR stages applied one after the
    // other to every window BEFORE the notch; stage outputs ping-pong between two chunk buffers
    std::vector<std::unique_ptr<FirBank>> prefilters;
    DevBuf d_pre[2];
    DevBuf d_nan_ptr, d_nan_cols;
    bool has_nan_cols = false;

    // families
    bool has_scan = false;
    int scan_h = 0, scan_r = 0, scan_l = 0;
    DevBuf d_scan_colmap;
    std::vector<std::unique_ptr<SpectralFam>> spectral;
    std::unique_ptr<BandpowerFam> bandpower;
    std::unique_ptr<BurstsFam> bursts;
    std::unique_ptr<SharpwaveFam> sharpwave;
    std::unique_ptr<NormFam> norm;
    std::unique_ptr<RawNormFam> rawnorm;

    // data
    DevBuf d_raw, d_xr, d_nanblk, d_gsum;
    bool raw_f64 = false;
    long long T = 0, raw_pitch = 0, xr_pitch = 0, nanblk_pitch = 0, gsum_pitch = 0;
    bool have_data = false, upload_pending = false, resident_uses_gsum = false;
    DevBuf d_starts, d_yoff, d_y, d_out, d_nanflags;
    long long out_rows = 0;
    long long out_pitch = 0;  // row pitch (elements) of the HOST result matrix; 0 = n_features (dense)
    int chunk = 1, Wp = 0;
    // arithmetic of the LINEAR families (notch, band power): 0 float64 (default), 1 float32 inside the FFT convolution.  Rows that
    // feed threshold / peak decisions (bursts, sharp waves, raw normaliser clip) always stay float64, and so does the notch then.
    int precision = 0;  // 0 float64, 1 scalar float32, 2 packed float32 pairs
    int f32_linear() const { return (precision != 0 && !bursts && !sharpwave && !rawnorm) ? precision : 0; }
    std::vector<long long> h_starts_one;
    DevBuf d_win_in;  // staging of a single streamed window
    std::unique_ptr<NmStream> strm;  // streaming entry (nm_stream.cuh): page-locked slot ring, parameter blocks, CUDA graphs

    int grid_for(size_t smem, int n_items, int threads) const {
        int occ = (int)std::max<size_t>(1, std::min<size_t>(2048 / threads, (size_t)(smem_max + 1024) / (smem + 1024)));
        long long g = (long long)n_sm * occ;
        return (int)std::max<long long>(1, std::min<long long>(g, n_items));
    }
};

#define NM_P_CHECK(p) NM_CHECK((p) != nullptr, "pipeline is NULL")

// stream operations that a replayed CUDA graph already contains are skipped while the host pass only patches that graph
#define NM_STREAM_OP(expr)                        \
    do {                                          \
        if (!nm_gs_updating()) NM_CUDA_CHECK(expr); \
    } while (0)

// Small per-launch host arrays (quantile positions, history ranges, ...) -> device.
//   batched paths:   DevBuf upload from the (temporary) vector; *need_sync tells the caller to synchronise before the vector dies
//   streaming pass:  the values are written into the slot's page-locked parameter block and copied by a small H2D on the SAME
//                    stream ahead of the kernel that reads them -- fixed addresses (graph replay), no synchronisation
template <typename T>
static int nm_param_upload(nm_pipeline* p, DevBuf& buf, const std::vector<T>& v, const T** dev, bool* need_sync) {
    NmStream* st = p->strm.get();
    if (st && st->cur >= 0) {
        const size_t bytes = v.size() * sizeof(T);
        const size_t off = (st->par_used + 15) & ~(size_t)15;
        NM_CHECK(off + bytes <= NM_STREAM_PAR_BYTES, "streaming parameter block too small");
        unsigned char* h = st->slots[st->cur].par_host + off;
        memcpy(h, v.data(), bytes);
        unsigned char* d = st->d_par.as<unsigned char>() + off;
        st->par_used = off + bytes;
        if (bytes) NM_STREAM_OP(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, p->stream));
        *dev = reinterpret_cast<const T*>(d);
        return 0;
    }
    if (buf.upload(v, p->stream)) return -1;
    *dev = buf.as<T>();
    if (need_sync) *need_sync = true;
    return 0;
}

template <class K>
static int nm_allow_smem(K kernel, size_t bytes, const nm_pipeline* p) {
    NM_CHECK((long long)bytes <= (long long)p->smem_max, "kernel needs %zu bytes of shared memory (> %d available per CTA)", bytes,
             p->smem_max);
    if (bytes > 48 * 1024) NM_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

// nm_prep_kernel over samples [a.t0, a.t1), specialised when there is at most one reference group
static void nm_launch_prep(nm_pipeline* p, const NmPrepArgs& a) {
    const unsigned grid = (unsigned)((a.t1 - a.t0 + 31) / 32);
    p->prof_begin();
    if (a.G <= 1 && a.raw_is_f64) {
        NM_LAUNCH((nm_prep_kernel<1, 1>), dim3(grid), dim3(NM_PREP_THREADS), nm_prep_smem_bytes(), p->stream, a);
    } else if (a.G <= 1) {
        NM_LAUNCH((nm_prep_kernel<1, 0>), dim3(grid), dim3(NM_PREP_THREADS), nm_prep_smem_bytes(), p->stream, a);
    } else {
        NM_LAUNCH((nm_prep_kernel<NM_MAX_GROUPS, -1>), dim3(grid), dim3(NM_PREP_THREADS), nm_prep_smem_bytes(), p->stream, a);
    }
    p->prof_end(NM_PROF_PREP);
    p->launches++;
}

// ------------------------------------------------------------------------------- FIR launches
// Three kernels implement the same contract; the most specialised one that covers the bank is used:
//   nm_convx_kernel  P in {1024, 2048, 4096}, compile-time plan (nm_convx.cuh)   <- every default configuration
//   nm_conv_kernel   other powers of two in [512, 8192], runtime plan (nm_conv.cuh)
//   nm_fir_kernel    5-smooth sizes, generic mixed radix (nm_fir.cuh)
template <class Epi>
using NmConvKernel = void (*)(NmConvArgs, Epi);

template <typename T, class Epi>
static NmConvKernel<Epi> nm_convx_pick_t(const FirBank& b) {
    if (b.mode == NM_FIR_REFLECT) {
        if constexpr (Epi::kReflectOk) {
            if (b.nF != 1 || b.mixed) return nullptr;
            switch (b.P) {
                case 1024: return nm_convx_kernel<T, 1024, true, false, Epi>;
                case 2048: return nm_convx_kernel<T, 2048, true, false, Epi>;
                default: return nm_convx_kernel<T, 4096, true, false, Epi>;
            }
        }
    } else {
        if constexpr (Epi::kSameOk) {
            switch (b.P) {
                case 1024: return nm_convx_kernel<T, 1024, false, true, Epi>;
                case 1536: return nm_convx_kernel<T, 1536, false, true, Epi>;
                case 2048: return nm_convx_kernel<T, 2048, false, true, Epi>;
                case 3072: return nm_convx_kernel<T, 3072, false, true, Epi>;
                default: return nm_convx_kernel<T, 4096, false, true, Epi>;
            }
        }
    }
    return nullptr;
}

// f32: float32 arithmetic inside the transform (optional fast mode; only epilogues that work from registers offer it):
// 1 = scalar float32, 2 = packed float32 pairs (two channel pairs per item, FFMA2 / FADD2 / FMUL2)
template <class Epi>
static NmConvKernel<Epi> nm_convx_pick(const FirBank& b, int f32 = 0) {
    if (!b.pow2 || !nm_convx_supported(b.P)) return nullptr;
    if constexpr (Epi::kF32Ok) {
        if (f32 == 2) return nm_convx_pick_t<f32x2, Epi>(b);
        if (f32 == 1) return nm_convx_pick_t<float, Epi>(b);
    }
    return nm_convx_pick_t<double, Epi>(b);
}

// 'same'-mode bank run filter by filter on the single-buffer instantiation (BANK = false): one more forward transform per extra
// filter, but half the shared memory per CTA -> more resident CTAs for the latency-bound shared-memory epilogues (sharp waves,
// burst envelopes).  Chosen per family at nm_finalize (split_filters); NMB200_SPLIT_BANKS=0 disables it.
template <class Epi>
static NmConvKernel<Epi> nm_convx_pick_single(const FirBank& b) {
    if (!b.pow2 || !nm_convx_supported(b.P) || b.mode != NM_FIR_SAME) return nullptr;
    if constexpr (Epi::kSameOk && Epi::kSplitOk) {
        switch (b.P) {
            case 1024: return nm_convx_kernel<double, 1024, false, false, Epi>;
            case 1536: return nm_convx_kernel<double, 1536, false, false, Epi>;
            case 2048: return nm_convx_kernel<double, 2048, false, false, Epi>;
            case 3072: return nm_convx_kernel<double, 3072, false, false, Epi>;
            default: return nm_convx_kernel<double, 4096, false, false, Epi>;
        }
    }
    return nullptr;
}
static bool nm_split_banks_enabled() {
    const char* env = getenv("NMB200_SPLIT_BANKS");
    return env ? atoi(env) != 0 : true;
}

// a bank whose spectrum + work buffers do not fit the shared memory of one CTA (long filters at high sampling rates: P = 8192
// and more than one filter) runs filter by filter: the forward transform is repeated, the buffers are not
template <class Epi>
static bool nm_fir_split(const FirBank& bank, size_t epi_bytes, const nm_pipeline* p) {
    if (nm_convx_pick<Epi>(bank)) return false;
    return bank.nF > 1 && bank.smem(epi_bytes) > (size_t)p->smem_max && bank.smem(epi_bytes, 1) <= (size_t)p->smem_max;
}

// dynamic shared memory of the kernel that nm_launch_fir will pick for (bank, Epi)
template <class Epi>
static size_t nm_fir_smem(const FirBank& bank, size_t epi_bytes) {
    if (nm_convx_pick<Epi>(bank)) return bank.smem_x(epi_bytes);
    return bank.smem(epi_bytes);
}

template <class Epi>
static int nm_allow_fir_smem(const FirBank& bank, size_t epi_bytes, const nm_pipeline* p, int f32 = 0) {
    if constexpr (Epi::kSplitOk) {
        if (auto k1 = nm_convx_pick_single<Epi>(bank)) {
            if (nm_allow_smem(k1, bank.smem_x(epi_bytes, 0, true), p)) return -1;
        }
    }
    if (auto k = nm_convx_pick<Epi>(bank, f32)) return nm_allow_smem(k, bank.smem_x(epi_bytes, Epi::kF32Ok ? f32 : 0), p);
    if constexpr (!Epi::kConvxOnly) {
        const size_t sm = bank.smem(epi_bytes, nm_fir_split<Epi>(bank, epi_bytes, p) ? 1 : -1);
        if (bank.pow2) return nm_allow_smem(nm_conv_kernel<Epi>, sm, p);
        return nm_allow_smem(nm_fir_kernel<Epi>, sm, p);
    }
    return 0;
}

template <class K>
static int nm_resident_grid(const nm_pipeline* p, K kernel, int threads, size_t smem_bytes, int n_items) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem_bytes) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long long g = (long long)p->n_sm * per_sm;  // persistent: exactly one resident wave, items strided over it
    return (int)std::max<long long>(1, std::min<long long>(g, n_items));
}

template <class Epi>
static void nm_launch_fir(nm_pipeline* p, const FirBank& bank, const NmRows& rows, const Epi& epi, cudaStream_t stream, size_t epi_bytes,
                          int f32 = 0) {
    const int threads = bank.threads();
    if constexpr (Epi::kSplitOk) {
        if (auto k1 = (bank.nF > 1 && nm_split_banks_enabled()) ? nm_convx_pick_single<Epi>(bank) : nullptr) {
            const size_t sm = bank.smem_x(epi_bytes, 0, true);
            for (int f = 0; f < bank.nF; ++f) {
                NmConvArgs a = bank.conv_args(rows, f, 1);
                a.scratch_in_tail = bank.epi_fits_tail(epi_bytes) ? 1 : 0;
                const int grid = nm_resident_grid(p, k1, threads, sm, a.n_items);
                NM_LAUNCH(k1, dim3(grid), dim3(threads), sm, stream, a, epi);
            }
            return;
        }
    }
    if (auto k = nm_convx_pick<Epi>(bank, f32)) {
        NmConvArgs a = bank.conv_args(rows);
        a.scratch_in_tail = bank.epi_fits_tail(epi_bytes) ? 1 : 0;
        if (f32 == 2 && Epi::kF32Ok) a.n_items = rows.n_windows * ((rows.n_ch + 3) / 4);  // channel quads
        const size_t sm = bank.smem_x(epi_bytes, Epi::kF32Ok ? f32 : 0);
        const int grid = nm_resident_grid(p, k, threads, sm, a.n_items);
        NM_LAUNCH(k, dim3(grid), dim3(threads), sm, stream, a, epi);
        return;
    }
    if constexpr (Epi::kConvxOnly) return;
    else {
        const bool split = nm_fir_split<Epi>(bank, epi_bytes, p);
        const int n_launch = split ? bank.nF : 1, per = split ? 1 : bank.nF;
        for (int l = 0; l < n_launch; ++l) {
            const size_t sm = bank.smem(epi_bytes, per);
            if (bank.pow2) {
                NmConvArgs a = bank.conv_args(rows, l * per, per);
                a.scratch_in_tail = bank.epi_fits_tail(epi_bytes) ? 1 : 0;
                const int grid = nm_resident_grid(p, nm_conv_kernel<Epi>, threads, sm, a.n_items);
                NM_LAUNCH(nm_conv_kernel<Epi>, dim3(grid), dim3(threads), sm, stream, a, epi);
            } else {
                NmFirArgs a = bank.args(rows, l * per, per);
                const int grid = nm_resident_grid(p, nm_fir_kernel<Epi>, threads, sm, a.n_items);
                NM_LAUNCH(nm_fir_kernel<Epi>, dim3(grid), dim3(threads), sm, stream, a, epi);
            }
        }
    }
}

// Notch of one batch of windows: rows -> y (and, on the specialised kernel, the scan features of the notched rows).
// Returns true when the scan family has been computed by the notch kernel's epilogue.
typedef void (*NmNotchxKernel)(NmConvArgs, NmEpiStoreScan);
static NmNotchxKernel nm_notchx_pick(int P) {
    return P == 1024 ? nm_notchx_kernel<1024> : (P == 2048 ? nm_notchx_kernel<2048> : nm_notchx_kernel<4096>);
}
static size_t nm_notchx_smem(int P, int W) {
    return P == 1024 ? nm_notchx_smem_bytes<1024>(W) : (P == 2048 ? nm_notchx_smem_bytes<2048>(W) : nm_notchx_smem_bytes<4096>(W));
}

static bool nm_launch_notch(nm_pipeline* p, const NmRows& rows, double* y, const NmOut* scan_out, cudaStream_t stream) {
    const FirBank& bank = *p->notch;
    if (nm_convx_pick<NmEpiStoreScan>(bank)) {
        NmEpiStoreScan epi;
        epi.y = y;
        epi.Wp = p->Wpi;
        epi.want_scan = scan_out ? 1 : 0;
        epi.want_hjorth = p->scan_h; epi.want_raw = p->scan_r; epi.want_ll = p->scan_l;
        if (scan_out) epi.out = *scan_out;
        else epi.out = NmOut{nullptr, 0, 0, nullptr, 0};
        if (p->notch_tma && !p->f32_linear()) {
            // float64 rows of the next item staged by the TMA engine (nm_notchx_kernel); row bases are 16-byte aligned: cudaMalloc'ed
            // buffers with even row pitches (xr_pitch, Wpi)
            NmConvArgs a = bank.conv_args(rows);
            auto k = nm_notchx_pick(bank.P);
            const size_t sm = nm_notchx_smem(bank.P, rows.W);
            const int grid = nm_resident_grid(p, k, bank.threads(), sm, a.n_items);
            NM_LAUNCH(k, dim3(grid), dim3(bank.threads()), sm, stream, a, epi);
            return scan_out != nullptr;
        }
        nm_launch_fir(p, bank, rows, epi, stream, 0, p->f32_linear());
        return scan_out != nullptr;
    }
    NmEpiStore epi{y, (long long)p->Wpi, 1};
    nm_launch_fir(p, bank, rows, epi, stream, 0);
    return false;
}

// ------------------------------------------------------------------------------- family launches
static NmOut nm_out_for(nm_pipeline* p, const DevBuf& colmap, int per_ch, int w0) {
    NmOut o;
    o.out = p->d_out.as<double>();
    o.row0 = w0;
    o.F = p->F;
    o.colmap = colmap.as<int>();
    o.per_ch = per_ch;
    return o;
}

int BurstsFam::allow_smem(const nm_pipeline* p) {
    if (nm_allow_fir_smem<NmEpiBursts>(bank, epi_smem(), p)) return -1;
    return nm_allow_smem(nm_burst_thr_kernel, thr_smem(), p);
}

// numpy 'linear' quantile bookkeeping (numpy/lib/_function_base_impl.py _quantile, alpha = beta = 1) for ALL windows of a
// run, uploaded once: the chunk launches then only offset into these arrays (no host synchronisation per chunk)
int BurstsFam::prepare(nm_pipeline* p, int n) {
    std::vector<long long> e_end(n);
    std::vector<int> nh(n), klo(n), khi(n);
    std::vector<double> gam(n);
    for (int k = 0; k < n; ++k) {
        const long long gw = batch + k;
        const long long e = (long long)W + gw * S;
        const int cnt = (int)std::min<long long>(ring_n, e);
        e_end[k] = e;
        nh[k] = cnt;
        const double virt = (double)cnt * q + (1.0 + q * (1.0 - 1.0 - 1.0)) - 1.0;
        long long lo = (long long)std::floor(virt), hi = lo + 1;
        double g = virt - (double)lo;
        if (virt >= (double)(cnt - 1)) { lo = hi = cnt - 1; }
        if (virt < 0) { lo = hi = 0; }
        klo[k] = (int)lo; khi[k] = (int)hi; gam[k] = g;
    }
    bool sync = false;
    if (nm_param_upload(p, d_e_end, e_end, &a_e_end, &sync) || nm_param_upload(p, d_n, nh, &a_n, &sync) ||
        nm_param_upload(p, d_lo, klo, &a_lo, &sync) || nm_param_upload(p, d_hi, khi, &a_hi, &sync) ||
        nm_param_upload(p, d_gamma, gam, &a_gamma, &sync))
        return -1;
    if (sync) NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));  // the staging vectors above are temporaries
    run_base = batch;
    return 0;
}

int BurstsFam::run(nm_pipeline* p, const NmRows& rows, int w0) {
    const int n = rows.n_windows;
    const long long k0 = batch - run_base;  // offset of this chunk inside the arrays prepared for the run
    (void)w0;

    NmEpiBursts epi;
    epi.hfft = hfft.dev();
    epi.need_scratch = hfft.generic ? 1 : 0;
    epi.fast_n = fast_n();
    epi.env = d_env.as<double>();
    epi.Wp = Wp;
    epi.nB = nB;
    epi.ring = d_ring.as<double>();
    epi.cap = cap;
    epi.win0 = batch;
    epi.S = S;
    p->prof_begin();
    nm_launch_fir(p, bank, rows, epi, p->stream, epi_smem());
    p->prof_end(NM_PROF_BURST_ENV);

    NmBurstThrArgs ta;
    ta.ring = d_ring.as<double>();
    ta.cap = cap;
    ta.n_ch = C; ta.nB = nB; ta.n_windows = n;
    ta.e_end = a_e_end + k0;
    ta.n_hist = a_n + k0;
    ta.k_lo = a_lo + k0; ta.k_hi = a_hi + k0;
    ta.gamma = a_gamma + k0;
    ta.thr = d_thr.as<double>();
    ta.incremental = incremental;
    ta.n_split = incremental ? nm_burst_thr_split(C * nB, n, p->n_sm) : 1;
    qstate.bind(ta, p->stream);
    const int n_thr = n * C * nB;
    p->prof_begin();
    NM_LAUNCH(nm_burst_thr_kernel, dim3(C * nB * ta.n_split), dim3(NM_BQ_THREADS), thr_smem(), p->stream, ta);
    p->prof_end(NM_PROF_BURST_THR);

    NmBurstFeatArgs ba;
    ba.env = d_env.as<double>();
    ba.Wp = Wp;
    ba.thr = d_thr.as<double>();
    ba.n_windows = n; ba.n_ch = C; ba.nB = nB; ba.W = W;
    ba.sfreq = sfreq; ba.seg_s = seg_s;
    ba.out = nm_out_for(p, d_colmap, nB * 6, w0);
    const int wpc = NM_ROW_THREADS / 32;
    p->prof_begin();
    NM_LAUNCH(nm_burst_feat_kernel, dim3(std::max(1, std::min((n_thr + wpc - 1) / wpc, p->n_sm * 16))), dim3(NM_ROW_THREADS), 0, p->stream, ba);
    p->prof_end(NM_PROF_BURST_FEAT);
    p->launches += 3;
    batch += n;
    return 0;
}

int SharpwaveFam::allow_smem(const nm_pipeline* p) { return nm_allow_fir_smem<NmEpiSharpwave>(bank, epi_smem(), p); }

int SharpwaveFam::run(nm_pipeline* p, const NmRows& rows, int w0) {
    NmEpiSharpwave epi;
    epi.cfg = cfg;
    epi.out = nm_out_for(p, d_colmap, per_ch, w0);
    p->prof_begin();
    nm_launch_fir(p, bank, rows, epi, p->stream, epi_smem());
    p->prof_end(NM_PROF_SHARPWAVE);
    p->launches++;
    return 0;
}

int RawNormFam::run(nm_pipeline* p, NmRows& rows) {
    const int n = rows.n_windows;
    // history bookkeeping per window (processing/normalization.py:92-107): statistics range [lo, end of block g)
    std::vector<long long> lo(n, 0), e_end(n, 0);
    const int nt = n_tracks();
    std::vector<int> nh(n, 0), klo[3], khi[3];
    std::vector<double> gam[3];
    for (int t = 0; t < nt; ++t) { klo[t].assign(n, 0); khi[t].assign(n, 0); gam[t].assign(n, 0.0); }
    for (int k = 0; k < n; ++k) {
        const long long g = batch + k;
        long long hist = W;
        e_end[k] = (long long)W + g * add;
        if (g == 0) {
            len_prev = W;
        } else {
            hist = len_prev + add;
            lo[k] = e_end[k] - hist;
            len_prev = (n_keep > 1) ? std::min<long long>(hist, n_keep - 1) : hist;  // previous[-n_keep + 1:] keeps everything for n_keep == 1
        }
        // numpy 'linear' quantile: virtual index (hist - 1) * q, the order statistics below / above it and the lerp weight
        // (q = 0.5: numpy.median, the two middle values averaged)
        nh[k] = (int)hist;
        for (int t = 0; t < nt; ++t) {
            const double q = track_q(t);
            if (q == 0.5) {
                klo[t][k] = (int)((hist - 1) / 2);
                khi[t][k] = (int)(hist / 2);
                gam[t][k] = (hist % 2 == 0) ? 0.5 : 0.0;
            } else {
                const double vi = (double)(hist - 1) * q;
                const long long f = (long long)std::floor(vi);
                klo[t][k] = (int)f;
                khi[t][k] = (int)std::min<long long>(f + 1, hist - 1);
                gam[t][k] = vi - (double)f;
                if (gam[t][k] == 0.0) khi[t][k] = klo[t][k];
            }
        }
    }
    NM_CHECK(n_keep > 1 || (long long)W + (batch + n) * add < cap, "raw normalisation history exceeds the ring (normalization_time_s * sfreq == 1)");
    const long long* a_lo = nullptr;
    bool sync = false;
    if (nm_param_upload(p, d_lo, lo, &a_lo, &sync)) return -1;
    if (sync) NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));  // `lo` is a temporary
    NmRawNormArgs a;
    a.in = rows;
    a.out = d_out.as<double>();
    a.Wp = Wp;
    a.ring = d_ring.as<double>();
    a.cap = cap;
    a.blk = d_blk.as<double>();
    a.blk_cap = blk_cap;
    a.g0 = batch;
    a.add = add;
    a.lo = a_lo;
    a.method = method;
    a.clip = clip;
    a.med = nt > 0 ? d_med[0].as<double>() : nullptr;
    a.q1 = nt > 1 ? d_med[1].as<double>() : nullptr;
    a.q2 = nt > 2 ? d_med[2].as<double>() : nullptr;
    const int wpc = NM_ROW_THREADS / 32;
    const long long n_rows = (long long)n * C;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_rows + wpc - 1) / wpc, (long long)p->n_sm * 16));
    p->prof_begin();
    NM_LAUNCH(nm_rawnorm_append_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, a);
    if (nt > 0) {
        const long long* m_e_end = nullptr;
        const int* m_n = nullptr;
        bool msync = false;
        if (nm_param_upload(p, d_e_end, e_end, &m_e_end, &msync) || nm_param_upload(p, d_n, nh, &m_n, &msync)) return -1;
        const int *m_lo[3] = {nullptr, nullptr, nullptr}, *m_hi[3] = {nullptr, nullptr, nullptr};
        const double* m_gamma[3] = {nullptr, nullptr, nullptr};
        for (int t = 0; t < nt; ++t)
            if (nm_param_upload(p, d_klo[t], klo[t], &m_lo[t], &msync) || nm_param_upload(p, d_khi[t], khi[t], &m_hi[t], &msync) ||
                nm_param_upload(p, d_gamma[t], gam[t], &m_gamma[t], &msync))
                return -1;
        if (msync) NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
        for (int t = 0; t < nt; ++t) {
            NmBurstThrArgs ta;
            ta.ring = d_ring.as<double>();
            ta.cap = cap;
            ta.n_ch = C; ta.nB = 1; ta.n_windows = n;
            ta.e_end = m_e_end;
            ta.n_hist = m_n;
            ta.k_lo = m_lo[t]; ta.k_hi = m_hi[t];
            ta.gamma = m_gamma[t];
            ta.thr = d_med[t].as<double>();
            ta.incremental = 1;
            ta.n_split = nm_burst_thr_split(C, n, p->n_sm);
            qstate[t].bind(ta, p->stream);
            NM_LAUNCH(nm_burst_thr_kernel, dim3(C * ta.n_split), dim3(NM_BQ_THREADS), BurstsFam::thr_smem(), p->stream, ta);
            p->launches++;
        }
    }
    NM_LAUNCH(nm_rawnorm_apply_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, a);
    p->prof_end(NM_PROF_NOTCH);
    p->launches += 2;
    batch += n;
    rows.base = d_out.as<double>();
    rows.ch_stride = Wp;
    rows.off = (p->has_resampler ? p->d_roff : p->d_yoff).as<long long>();
    return 0;
}


int NormFam::run(nm_pipeline* p, int n_windows, int w0, int total) {
    if (n_cols == 0) return 0;
    // History of RAW rows: a fixed-capacity block of cap = n_keep - 1 rows, the valid ones RIGHT-aligned.  The streaming entry
    // (one window per call) then always copies cap rows in and out -- constant sizes, so its CUDA graph never changes shape --
    // and passes n_prev = cap: the kernel reads rows [row - (nh - 1), row] with nh <= g + 1, i.e. never an unwritten one.
    // Chunked (total > 0): rows [w0, w0 + n_windows) of a run of `total` windows; the history is copied in with the first chunk and
    // out with the last one, the raw rows of the run accumulate in `ext` in between.
    const int cap = std::max(0, n_keep - 1);
    const bool streaming = p->strm && p->strm->cur >= 0 && n_windows == 1;
    const bool chunked = total > 0;
    const int run_windows = chunked ? total : n_windows;
    const int prev = streaming ? cap : n_prev;
    const size_t rows = (size_t)prev + run_windows;
    if (!chunked || w0 == 0) {
        if (d_ext.ensure(std::max(rows, (size_t)cap + 1) * n_cols * sizeof(double))) return -1;
    }
    double* ext = d_ext.as<double>();
    const double* hist_valid = d_hist.as<double>() + (size_t)(cap - prev) * n_cols;
    if (prev && (!chunked || w0 == 0))
        NM_STREAM_OP(cudaMemcpyAsync(ext, hist_valid, (size_t)prev * n_cols * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    const int r0 = chunked ? w0 : 0;  // first row of this call behind the history block
    const long long tot = (long long)n_windows * n_cols;
    const unsigned grid = (unsigned)((tot + NM_ROW_THREADS - 1) / NM_ROW_THREADS);
    double* const out_rows = p->d_out.as<double>() + (size_t)w0 * p->F;
    NM_LAUNCH(nm_norm_gather_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, (const double*)out_rows, p->F,
              (const int*)d_cols.as<int>(), n_cols, n_windows, ext + ((size_t)prev + r0) * n_cols);
    NmNormArgs a;
    a.ext = ext;
    a.n_prev = prev + r0;
    a.n_windows = n_windows;
    a.n_cols = n_cols;
    a.cols = d_cols.as<int>();
    a.g0 = batch + r0;
    a.n_keep = n_keep;
    a.method = method;
    a.clip = clip;
    a.out = out_rows;
    a.F = p->F;
    p->prof_begin();
    // order-statistic methods over a batch of windows: the sliding sorted history (one thread per column); else one thread per
    // (window, column)
    int order_threads = 0;
    if ((method == 1 || method == 3 || method == 5) && n_windows > 1) {
        for (int tpb : {64, 32})
            if (!order_threads && (size_t)n_keep * tpb * sizeof(double) <= (size_t)p->smem_max) order_threads = tpb;
    }
    if (order_threads) {
        NmNormOrderArgs o;
        o.ext = ext; o.n_prev = a.n_prev; o.n_windows = n_windows; o.n_cols = n_cols; o.cols = d_cols.as<int>();
        o.g0 = a.g0; o.n_keep = n_keep; o.method = method; o.clip = clip; o.out = out_rows; o.F = p->F;
        const size_t sm = (size_t)n_keep * order_threads * sizeof(double);
        if (nm_allow_smem(nm_norm_order_kernel, sm, p)) return -1;
        NM_LAUNCH(nm_norm_order_kernel, dim3((n_cols + order_threads - 1) / order_threads), dim3(order_threads), sm, p->stream, o);
    } else {
        NM_LAUNCH(nm_norm_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, a);
    }
    p->prof_end(NM_PROF_NORM);
    p->launches += 2;
    if (chunked && w0 + n_windows < total) return 0;
    // keep the last cap raw rows for the next call (right-aligned)
    const int keep = streaming ? cap : (int)std::min<size_t>(rows, (size_t)cap);
    if (keep)
        NM_STREAM_OP(cudaMemcpyAsync(d_hist.as<double>() + (size_t)(cap - keep) * n_cols, ext + (rows - keep) * n_cols,
                                     (size_t)keep * n_cols * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    n_prev = (int)std::min<size_t>((size_t)n_prev + run_windows, (size_t)cap);
    batch += run_windows;
    return 0;
}


// ------------------------------------------------------------------------------- fused window kernel (nm_fused.cuh)
typedef void (*NmFusedKernel)(NmFusedArgs);
static NmFusedKernel nm_fused_pick(int P, int sx, bool raw64) {
    if (P == 1024) return raw64 ? nm_fused_kernel<1024, NmSxNone, true> : nm_fused_kernel<1024, NmSxNone, false>;
    if (P == 2048) {
        if (sx == 1000) return raw64 ? nm_fused_kernel<2048, NmSx1000, true> : nm_fused_kernel<2048, NmSx1000, false>;
        return raw64 ? nm_fused_kernel<2048, NmSxNone, true> : nm_fused_kernel<2048, NmSxNone, false>;
    }
    if (P == 4096) {
        if (sx == 2000) return raw64 ? nm_fused_kernel<4096, NmSx2000, true> : nm_fused_kernel<4096, NmSx2000, false>;
        return raw64 ? nm_fused_kernel<4096, NmSxNone, true> : nm_fused_kernel<4096, NmSxNone, false>;
    }
    return nullptr;
}
static size_t nm_fused_smem(int P) {
    return P == 1024 ? nm_fused_smem_bytes<1024>() : (P == 2048 ? nm_fused_smem_bytes<2048>() : nm_fused_smem_bytes<4096>());
}
static NmFusedKernel nm_front_pick(int P, int sx, bool raw64) {
    if (P == 1024) return raw64 ? nm_front_kernel<1024, NmSxNone, true> : nm_front_kernel<1024, NmSxNone, false>;
    if (P == 2048) {
        if (sx == 1000) return raw64 ? nm_front_kernel<2048, NmSx1000, true> : nm_front_kernel<2048, NmSx1000, false>;
        return raw64 ? nm_front_kernel<2048, NmSxNone, true> : nm_front_kernel<2048, NmSxNone, false>;
    }
    if (P == 4096) {
        if (sx == 2000) return raw64 ? nm_front_kernel<4096, NmSx2000, true> : nm_front_kernel<4096, NmSx2000, false>;
        return raw64 ? nm_front_kernel<4096, NmSxNone, true> : nm_front_kernel<4096, NmSxNone, false>;
    }
    return nullptr;
}
static size_t nm_front_smem(int P, int W, bool raw64) {
    if (P == 1024) return raw64 ? nm_front_smem_bytes<1024, true>(W) : nm_front_smem_bytes<1024, false>(W);
    if (P == 2048) return raw64 ? nm_front_smem_bytes<2048, true>(W) : nm_front_smem_bytes<2048, false>(W);
    return raw64 ? nm_front_smem_bytes<4096, true>(W) : nm_front_smem_bytes<4096, false>(W);
}
static size_t nm_fused_nbuf(int P) { return P == 1024 ? NmCxPlan<1024>::NBUF : (P == 2048 ? NmCxPlan<2048>::NBUF : NmCxPlan<4096>::NBUF); }

// decide at nm_finalize whether the window chain runs in the fused kernel and which families it serves
static int nm_fused_plan(nm_pipeline* p) {
    p->fused = false;
    p->front = false;
    p->fused_bp = false;
    p->fused_sx = 0;
    p->fused_spec.clear();
    int want = p->fused_mode;
    if (want < 0) {
        const char* env = getenv("NMB200_FUSED");
        want = env ? atoi(env) : 0;  // staged kernels are the fastest organisation measured on B200
    }
    if (want == 0) return 0;
    if (!p->notch || !p->reref_foldable || !p->prefilters.empty() || p->rawnorm || p->has_resampler || p->f32_linear() || p->precision != 0) return 0;
    const FirBank& nb = *p->notch;
    if (!nb.pow2 || !nm_convx_supported(nb.P) || nb.nF != 1 || nb.mode != NM_FIR_REFLECT) return 0;
    const size_t buf_bytes = nm_fused_nbuf(nb.P) * sizeof(cx<double>);
    if (NmFzStage<true>::bytes(p->W) > buf_bytes) return 0;  // the stage of one item must fit a transform buffer
    const int nat_elems = (NmEpiStoreScan::phys(p->W - 1) + 2) & ~1;
    if ((size_t)nat_elems * sizeof(cx<double>) > buf_bytes) return 0;
    if (want == 2 ? nm_front_smem(nb.P, p->W, true) > (size_t)p->smem_max : nm_fused_smem(nb.P) > (size_t)p->smem_max) return 0;
    p->fused = true;
    p->front = want == 2;
    int sx_ok = nb.P == 2048 ? 1000 : (nb.P == 4096 ? 2000 : 0);
    if (p->front) {  // the segment DFTs run faster as a kernel of their own (11 CTAs per SM instead of 4): opt-in inside the front kernel
        const char* env = getenv("NMB200_FRONT_DFT");
        if (!env || atoi(env) == 0) sx_ok = 0;
    }
    for (size_t i = 0; i < p->spectral.size() && sx_ok; ++i) {
        const SpectralFam& f = *p->spectral[i];
        const nm_spectral_cfg& c = f.cfg;
        const size_t vals = (size_t)2 * f.nk * sizeof(double);
        // front kernel: natural-order window, DFT buffer and bin values share ONE transform buffer
        const size_t dft_buf = p->front ? (size_t)f.nbuf() * sizeof(cx<double>) : 0;
        if (f.fast && c.nper == sx_ok && c.nseg == 1 && !c.keep_segments && c.start >= 0 && c.start + c.nper <= p->W &&
            (size_t)nat_elems * sizeof(cx<double>) + dft_buf + vals <= buf_bytes && (int)p->fused_spec.size() < NM_FZ_MAX_SPEC) {
            p->fused_spec.push_back((int)i);
            p->fused_sx = sx_ok;
        }
    }
    if (p->bandpower && !p->front) {
        const BandpowerFam& f = *p->bandpower;
        p->fused_bp = f.bank.pow2 && f.bank.P == nb.P && f.bank.mode == NM_FIR_SAME && !(f.mob || f.comp) && f.bank.d_hx.p != nullptr;
    }
    if (p->d_dfold.upload(p->h_dfold, p->stream) || p->d_gfold.upload(p->h_gfold, p->stream)) return -1;
    for (int raw64 = 0; raw64 < 2; ++raw64) {
        if (p->front) {
            if (nm_allow_smem(nm_front_pick(nb.P, p->fused_sx, raw64 != 0), nm_front_smem(nb.P, p->W, raw64 != 0), p)) return -1;
        } else if (nm_allow_smem(nm_fused_pick(nb.P, p->fused_sx, raw64 != 0), nm_fused_smem(nb.P), p)) {
            return -1;
        }
    }
    return 0;
}
static bool nm_spec_in_fused(const nm_pipeline* p, size_t i) {
    for (int k : p->fused_spec)
        if ((size_t)k == i) return true;
    return false;
}

static NmSpecArgs nm_spec_args(const nm_pipeline* p, const SpectralFam& f, const NmRows& rows, const NmOut& out, int n) {
    NmSpecArgs a;
    const nm_spectral_cfg& c = f.cfg;
    a.in = rows;
    a.fft = f.fft.dev();
    a.need_scratch = f.fft.generic ? 1 : 0;
    a.nseg = c.nseg; a.hop = c.hop; a.start = c.start;
    a.ext_even = c.ext_even; a.ext_len = c.ext_len;
    a.detrend = c.detrend;
    a.win = c.win;
    a.power = c.power; a.scale = c.scale; a.log = c.log;
    a.keep_segments = c.keep_segments;
    a.k0 = f.k0; a.nk = f.nk;
    a.n_bands = c.n_bands;
    a.band_lo = f.d_lo.as<int>(); a.band_hi = f.d_hi.as<int>();
    a.est_mask = c.est_mask;
    a.want_spectrum = c.want_spectrum;
    a.out = out;
    a.n_items = n * ((p->C + 1) / 2);
    return a;
}

// ------------------------------------------------------------------------------- life cycle
extern "C" int nm_pipeline_create(int device, int n_raw_rows, int n_ch, int window_samples, int n_features, nm_pipeline** out) {
    NM_CHECK(out, "out is NULL");
    NM_CHECK(n_raw_rows > 0 && n_ch > 0 && n_ch <= n_raw_rows, "bad channel counts (%d raw rows, %d feature channels)", n_raw_rows, n_ch);
    NM_CHECK(window_samples >= 3, "window must have at least 3 samples, got %d", window_samples);
    NM_CHECK(n_features > 0, "n_features must be positive");
    int n_dev = 0;
    NM_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
    NM_CHECK(n_dev > 0, "no CUDA device visible: libnmb200 has no CPU path");
    NM_CHECK(device >= 0 && device < n_dev, "device %d out of range (%d visible)", device, n_dev);
    NM_CUDA_CHECK(cudaSetDevice(device));
    auto p = std::make_unique<nm_pipeline>();
    p->device = device;
    p->C_all = n_raw_rows;
    p->C = n_ch;
    p->W = window_samples;
    p->Win = window_samples;
    p->F = n_features;
    NM_CUDA_CHECK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    NM_CUDA_CHECK(cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    NM_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_sync, cudaEventDisableTiming));
    NM_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    NM_CUDA_CHECK(cudaStreamCreateWithFlags(&p->red_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 3; ++b) {
        NM_CUDA_CHECK(cudaStreamCreateWithFlags(&p->side[b], cudaStreamNonBlocking));
        NM_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_join[b], cudaEventDisableTiming));
    }
    NM_CUDA_CHECK(cudaEventCreate(&p->ev0));
    NM_CUDA_CHECK(cudaEventCreate(&p->ev1));
    NM_CUDA_CHECK(cudaEventCreate(&p->pe0));
    NM_CUDA_CHECK(cudaEventCreate(&p->pe1));
    NM_CUDA_CHECK(cudaDeviceGetAttribute(&p->n_sm, cudaDevAttrMultiProcessorCount, device));
    NM_CUDA_CHECK(cudaDeviceGetAttribute(&p->smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    p->pick.resize(n_ch);
    for (int i = 0; i < n_ch; ++i) p->pick[i] = i;
    *out = p.release();
    return 0;
}

static void nm_stream_release(nm_pipeline* p);

static void nm_upload_join(nm_pipeline* p);
extern "C" void nm_pipeline_destroy(nm_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    nm_upload_join(p);
    if (p->out_stage) cudaFreeHost(p->out_stage);
    for (auto e : p->d2h_ev) cudaEventDestroy(e);
    for (int b = 0; b < 2; ++b) {
        if (p->up_stage[b]) cudaFreeHost(p->up_stage[b]);
        if (p->up_free[b]) cudaEventDestroy(p->up_free[b]);
    }
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->pe0) cudaEventDestroy(p->pe0);
    if (p->pe1) cudaEventDestroy(p->pe1);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    if (p->ev_sync) cudaEventDestroy(p->ev_sync);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->red_stream) { cudaStreamSynchronize(p->red_stream); cudaStreamDestroy(p->red_stream); }
    for (int b = 0; b < 3; ++b) {
        if (p->side[b]) { cudaStreamSynchronize(p->side[b]); cudaStreamDestroy(p->side[b]); }
        if (p->ev_join[b]) cudaEventDestroy(p->ev_join[b]);
    }
    for (auto e : p->slice_ev) cudaEventDestroy(e);
    for (auto e : p->red_ev) cudaEventDestroy(e);
    for (auto e : p->chunk_ev) cudaEventDestroy(e);
    nm_stream_release(p);
    cudaStream_t s = p->stream, cs = p->copy_stream;
    delete p;
    if (s) cudaStreamDestroy(s);
    if (cs) cudaStreamDestroy(cs);
}

extern "C" int nm_set_pick(nm_pipeline* p, const int* pick) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    for (int i = 0; i < p->C; ++i) {
        NM_CHECK(pick[i] >= 0 && pick[i] < p->C_all, "pick[%d] = %d out of range", i, pick[i]);
        p->pick[i] = pick[i];
    }
    return 0;
}

extern "C" int nm_set_reref(nm_pipeline* p, int n_groups, const int* group_of, const double* gcoef, const int* sp_ptr,
                            const int* sp_col, const double* sp_val) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(n_groups >= 0 && n_groups <= NM_MAX_GROUPS, "n_groups must be in [0, %d]", NM_MAX_GROUPS);
    NM_CHECK(sp_ptr && sp_ptr[0] == 0, "sp_ptr must start at 0");
    cudaSetDevice(p->device);
    const int nnz = sp_ptr[p->C];
    for (int k = 0; k < nnz; ++k) NM_CHECK(sp_col[k] >= 0 && sp_col[k] < p->C, "sp_col out of range");
    p->G = n_groups;
    std::vector<int> go(p->C, -1);
    if (n_groups > 0) {
        NM_CHECK(group_of && gcoef, "group arrays missing");
        for (int i = 0; i < p->C; ++i) {
            NM_CHECK(group_of[i] >= -1 && group_of[i] < n_groups, "group_of out of range");
            go[i] = group_of[i];
        }
    }
    if (p->d_group_of.upload(go, p->stream)) return -1;
    if (p->d_gcoef.upload(gcoef, (size_t)p->C * n_groups, p->stream)) return -1;
    if (p->d_sp_ptr.upload(sp_ptr, (size_t)p->C + 1, p->stream)) return -1;
    if (p->d_sp_col.upload(sp_col, (size_t)nnz, p->stream)) return -1;
    if (p->d_sp_val.upload(sp_val, (size_t)nnz, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->has_reref = true;
    // can the matrix be folded into the fused kernel's load?  row i = d_i on its own channel + g_i on the (single) group sum
    p->reref_foldable = n_groups <= 1;
    p->h_dfold.assign(p->C, 0.0);
    p->h_gfold.assign(p->C, 0.0);
    for (int i = 0; i < p->C && p->reref_foldable; ++i) {
        if (n_groups == 1) p->h_gfold[i] = gcoef[i];
        for (int k = sp_ptr[i]; k < sp_ptr[i + 1]; ++k) {
            if (sp_col[k] != i || k != sp_ptr[i]) { p->reref_foldable = false; break; }
            p->h_dfold[i] = sp_val[k];
        }
    }
    return 0;
}

extern "C" int nm_set_notch(nm_pipeline* p, const double* taps, int n_taps) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    if (n_taps <= 0) {
        p->notch.reset();
        return 0;
    }
    NM_CHECK(taps, "taps is NULL");
    cudaSetDevice(p->device);
    p->notch = std::make_unique<FirBank>();
    if (p->notch->build(taps, 1, n_taps, p->Win, NM_FIR_REFLECT, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

// MNE's padding for an n-sample row (npad="auto"): min(n // 8, 100) * 2 extra samples, then up to the next power of two
static int nm_resample_fast_plan(int n_in, int n_out, int D, int* P_out, int* pad_out) {
    if (D < 2 || n_in % 2 != 0 || n_in != n_out * D) return 0;
    const int min_add = std::min(n_in / 8, 100) * 2;
    int P = 1;
    while (P < n_in + min_add) P <<= 1;
    const int pad_each = (P - n_in) / 2;
    if (!nm_convx_supported(P) || (P - n_in) % 2 != 0 || pad_each % D != 0 || pad_each > n_in - 1 || P % (2 * D) != 0) return 0;
    *P_out = P;
    *pad_out = pad_each;
    return 1;
}

extern "C" int nm_set_resampler(nm_pipeline* p, int n_in, const double* op, int fft_decim) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(!p->notch && p->prefilters.empty(), "nm_set_resampler must be called before nm_add_prefilter / nm_set_notch (they filter the un-resampled window)");
    NM_CHECK(op && n_in >= 3, "bad resampling operator");
    cudaSetDevice(p->device);
    p->Win = n_in;
    p->rs_pitch = (p->W + 3) & ~3;
    p->rs_bank.reset();
    p->rs_decim = 0;
    int P = 0, pad_each = 0;
    if (fft_decim >= 2 && nm_resample_fast_plan(n_in, p->W, fft_decim, &P, &pad_each)) {
        // the caller says `op` is MNE's default FFT down-sampler by an integer factor: check a few interior entries of the operator
        // against the closed form of the fast path before trusting the hint (g = circular impulse response of the ideal low-pass)
        const int D = fft_decim, nyq = P / (2 * D);
        auto g = [&](long long n) {
            n = ((n % P) + P) % P;
            double acc = 1.0;
            for (int f = 1; f <= nyq; ++f) acc += 2.0 * std::cos(2.0 * M_PI * (double)((n * f) % P) / (double)P);
            return acc / (double)P;
        };
        bool ok = true;
        const int probes[4][2] = {{0, 1}, {p->W / 3, n_in / 2}, {p->W - 1, n_in - 2}, {p->W / 2, 3}};
        for (const auto& pr : probes) {
            const int j = pr[0], k = pr[1];
            if (k <= 0 || k >= n_in - 1 || j < 0 || j >= p->W) continue;
            const long long at = (long long)pad_each + (long long)D * j;  // padded position of output j
            double r = g(at - (pad_each + k));
            if (k <= pad_each) r -= g(at - (pad_each - k));                                  // left mirror: 2 x[0] - x[k]
            if (n_in - 1 - k <= pad_each && n_in - 1 - k >= 1) r -= g(at - (pad_each + n_in - 1 + (n_in - 1 - k)));  // right mirror
            if (std::fabs(r - op[(size_t)j * n_in + k]) > 1e-9) ok = false;
        }
        if (ok) {
            p->rs_bank = std::make_unique<FirBank>();
            if (p->rs_bank->build_lowpass(n_in, P, pad_each, D, p->stream)) return -1;
            p->rs_decim = D;
        }
    }
    if (!p->rs_bank) {
        std::vector<double> rt((size_t)n_in * p->rs_pitch, 0.0);  // transposed: the GEMM reads it k-major
        for (int j = 0; j < p->W; ++j)
            for (int k = 0; k < n_in; ++k) rt[(size_t)k * p->rs_pitch + j] = op[(size_t)j * n_in + k];
        if (p->d_rt.upload(rt, p->stream)) return -1;
    }
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->has_resampler = true;
    return 0;
}

extern "C" int nm_add_prefilter(nm_pipeline* p, const double* taps, int n_taps) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(taps && n_taps > 0, "bad prefilter taps");
    cudaSetDevice(p->device);
    auto b = std::make_unique<FirBank>();
    if (b->build(taps, 1, n_taps, p->Win, NM_FIR_SAME, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->prefilters.push_back(std::move(b));
    return 0;
}

extern "C" int nm_set_nan_columns(nm_pipeline* p, const int* col_ptr, const int* cols) {
    NM_P_CHECK(p);
    NM_CHECK(col_ptr && col_ptr[0] == 0, "col_ptr must start at 0");
    cudaSetDevice(p->device);
    const int n = col_ptr[p->C_all];
    for (int k = 0; k < n; ++k) NM_CHECK(cols[k] >= 0 && cols[k] < p->F, "NaN column out of range");
    if (p->d_nan_ptr.upload(col_ptr, (size_t)p->C_all + 1, p->stream)) return -1;
    if (p->d_nan_cols.upload(cols, (size_t)n, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->has_nan_cols = true;
    return 0;
}

static int nm_check_colmap(const nm_pipeline* p, const int* colmap, size_t n) {
    NM_CHECK(colmap, "colmap is NULL");
    for (size_t i = 0; i < n; ++i) NM_CHECK(colmap[i] >= -1 && colmap[i] < p->F, "colmap[%zu] = %d out of range (F = %d)", i, colmap[i], p->F);
    return 0;
}

extern "C" int nm_add_scan(nm_pipeline* p, int hjorth, int raw, int linelength, const int* colmap) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    if (nm_check_colmap(p, colmap, (size_t)p->C * 5)) return -1;
    cudaSetDevice(p->device);
    p->has_scan = true;
    p->scan_h = hjorth; p->scan_r = raw; p->scan_l = linelength;
    if (p->d_scan_colmap.upload(colmap, (size_t)p->C * 5, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int nm_add_spectral(nm_pipeline* p, const nm_spectral_cfg* cfg) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(cfg && cfg->nper >= 2 && cfg->nseg >= 1 && cfg->n_bands >= 0, "bad spectral configuration");
    cudaSetDevice(p->device);
    auto f = std::make_unique<SpectralFam>();
    f->cfg = *cfg;
    const int nbins = cfg->nper / 2 + 1;
    f->per_ch = cfg->n_bands * 4 + nbins;
    if (nm_check_colmap(p, cfg->colmap, (size_t)p->C * f->per_ch)) return -1;
    int kmin = nbins, kmax = 0;
    for (int b = 0; b < cfg->n_bands; ++b) {
        NM_CHECK(cfg->band_lo[b] >= 0 && cfg->band_hi[b] <= nbins && cfg->band_lo[b] <= cfg->band_hi[b],
                 "band %d bin range [%d, %d) outside the %d-bin spectrum", b, cfg->band_lo[b], cfg->band_hi[b], nbins);
        if (cfg->band_hi[b] > cfg->band_lo[b]) {
            kmin = std::min(kmin, cfg->band_lo[b]);
            kmax = std::max(kmax, cfg->band_hi[b]);
        }
    }
    if (cfg->want_spectrum) { kmin = 0; kmax = nbins; }
    if (kmax <= kmin) { kmin = 0; kmax = 1; }
    f->k0 = kmin;
    f->nk = kmax - kmin;
    f->fast = nm_specx_supported(cfg->nper);
    if (f->fast) {
        const std::vector<int> radices = cfg->nper == 1000 ? std::vector<int>{10, 10, 10}
                                         : (cfg->nper == 2000 ? std::vector<int>{20, 10, 10} : std::vector<int>{10, 10, 5});
        if (f->fft.build_with(cfg->nper, radices, p->stream)) return -1;
    } else if (f->fft.build(cfg->nper, p->stream)) {
        return -1;
    }
    if (cfg->win && f->d_win.upload(cfg->win, (size_t)cfg->nper, p->stream)) return -1;
    if (f->d_lo.upload(cfg->band_lo, (size_t)cfg->n_bands, p->stream)) return -1;
    if (f->d_hi.upload(cfg->band_hi, (size_t)cfg->n_bands, p->stream)) return -1;
    if (f->d_colmap.upload(cfg->colmap, (size_t)p->C * f->per_ch, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    f->cfg.win = cfg->win ? f->d_win.as<double>() : nullptr;
    p->spectral.push_back(std::move(f));
    return 0;
}

extern "C" int nm_add_bandpower(nm_pipeline* p, int n_bands, const double* taps, int n_taps, const int* seglen, int activity,
                                int mobility, int complexity, int log_transform, const int* colmap) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(n_bands > 0 && taps && n_taps > 0 && seglen, "bad band-power configuration");
    if (nm_check_colmap(p, colmap, (size_t)p->C * n_bands * 3)) return -1;
    cudaSetDevice(p->device);
    auto f = std::make_unique<BandpowerFam>();
    if (f->bank.build(taps, n_bands, n_taps, p->W, NM_FIR_SAME, p->stream)) return -1;
    for (int b = 0; b < n_bands; ++b) NM_CHECK(seglen[b] >= 1, "segment length must be >= 1 sample");
    if (f->d_seglen.upload(seglen, (size_t)n_bands, p->stream)) return -1;
    f->h_seglen.assign(seglen, seglen + n_bands);
    if (f->d_colmap.upload(colmap, (size_t)p->C * n_bands * 3, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    f->act = activity; f->mob = mobility; f->comp = complexity; f->logt = log_transform;
    p->bandpower = std::move(f);
    return 0;
}

extern "C" int nm_add_bursts(nm_pipeline* p, int n_bands, const double* taps, int n_taps, int samples_overlap, int ring_samples,
                             double quantile, double sfreq, double segment_length_s, const int* colmap) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(n_bands > 0 && taps && n_taps > 0, "bad bursts configuration");
    NM_CHECK(samples_overlap >= 0 && samples_overlap <= p->W, "samples_overlap must be in [0, W]");
    NM_CHECK(ring_samples >= 1, "ring_samples must be >= 1");
    NM_CHECK(quantile >= 0.0 && quantile <= 1.0, "quantile must be in [0, 1]");
    if (nm_check_colmap(p, colmap, (size_t)p->C * n_bands * 6)) return -1;
    cudaSetDevice(p->device);
    auto f = std::make_unique<BurstsFam>();
    if (f->build(taps, n_bands, n_taps, p->C, p->W, samples_overlap, ring_samples, quantile, sfreq, segment_length_s, colmap, p->stream))
        return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->bursts = std::move(f);
    return 0;
}

extern "C" int nm_add_sharpwave(nm_pipeline* p, int n_filters, const double* taps, int n_taps, int dist_peaks, int dist_troughs,
                                int sharp_offset, double ms_per_sample, int n_combo, const int* feat_ids, const int* est_ids,
                                int pair_estimator, int want_num_peaks, const int* colmap) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(n_filters > 0 && taps && n_taps > 0 && n_combo >= 0, "bad sharp-wave configuration");
    if (nm_check_colmap(p, colmap, (size_t)p->C * n_filters * (n_combo + 1))) return -1;
    cudaSetDevice(p->device);
    auto f = std::make_unique<SharpwaveFam>();
    if (f->build(taps, n_filters, n_taps, p->C, p->W, dist_peaks, dist_troughs, sharp_offset, ms_per_sample, n_combo, feat_ids, est_ids,
                 pair_estimator, want_num_peaks, colmap, p->stream))
        return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->sharpwave = std::move(f);
    return 0;
}

extern "C" int nm_add_feature_normalizer(nm_pipeline* p, int method, double clip, int n_keep, int n_cols, const int* cols) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(method >= 0 && method <= 6, "normalisation method must be 0..6 (mean, median, zscore, zscore-median, minmax, robust, quantile)");
    NM_CHECK(n_keep >= 1 && n_cols >= 0, "bad normaliser configuration");
    NM_CHECK(method != 6 || n_keep <= 300, "the quantile normaliser covers histories of at most 300 windows (n_quantiles = 300)");
    for (int i = 0; i < n_cols; ++i) NM_CHECK(cols[i] >= 0 && cols[i] < p->F, "normaliser column out of range");
    cudaSetDevice(p->device);
    auto f = std::make_unique<NormFam>();
    if (f->build(method, clip, n_keep, n_cols, cols, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->norm = std::move(f);
    return 0;
}

extern "C" int nm_set_precision(nm_pipeline* p, int float32_linear) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(float32_linear >= 0 && float32_linear <= 2, "precision must be 0 (float64), 1 (float32) or 2 (packed float32 pairs)");
    p->precision = float32_linear;
    return 0;
}

extern "C" int nm_set_fused(nm_pipeline* p, int mode) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(mode >= -1 && mode <= 2, "fused mode must be -1 (environment), 0 (staged kernels), 1 (one kernel) or 2 (front kernel)");
    p->fused_mode = mode;
    return 0;
}

extern "C" int nm_set_raw_normalizer(nm_pipeline* p, int method, double clip, int n_keep, int add_samples) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    NM_CHECK(method >= 0 && method <= 5, "raw normalisation method must be 0..5 (mean, median, zscore, zscore-median, minmax, robust); got %d", method);
    NM_CHECK(n_keep >= 1 && add_samples >= 1 && clip >= 0.0, "bad raw normaliser configuration");
    auto f = std::make_unique<RawNormFam>();
    if (f->build(method, clip, n_keep, add_samples, p->C, p->W)) return -1;
    p->rawnorm = std::move(f);
    return 0;
}

extern "C" int nm_finalize(nm_pipeline* p) {
    NM_P_CHECK(p);
    NM_CHECK(!p->finalized, "pipeline already finalized");
    cudaSetDevice(p->device);
    if (p->d_pick.upload(p->pick, p->stream)) return -1;
    if (!p->has_reref) {  // identity
        std::vector<int> go(p->C, -1), ptr(p->C + 1), col(p->C);
        std::vector<double> val(p->C, 1.0);
        for (int i = 0; i <= p->C; ++i) ptr[i] = i;
        for (int i = 0; i < p->C; ++i) col[i] = i;
        p->G = 0;
        if (p->d_group_of.upload(go, p->stream) || p->d_gcoef.ensure(16) || p->d_sp_ptr.upload(ptr, p->stream) ||
            p->d_sp_col.upload(col, p->stream) || p->d_sp_val.upload(val, p->stream))
            return -1;
    }
    if (!p->has_reref) {
        p->reref_foldable = true;
        p->h_dfold.assign(p->C, 1.0);
        p->h_gfold.assign(p->C, 0.0);
    }
    if (nm_fused_plan(p)) return -1;
    p->Wp = (p->W + 1) & ~1;
    p->Wpi = (p->Win + 1) & ~1;
    // chunk of windows per launch: as many as possible up to 64 windows / 384 MB of notched rows (+ burst envelopes).  L2 residency
    // of the chunk turned out not to matter (the consumers are compute / latency bound: a 24..128 MB sweep moved the C3 step by
    // < 3 %), while larger launches fill the persistent grids better: -11 % on the default feature set from 16 -> 64 windows
    const size_t per_window = (size_t)p->C * std::max(p->Wp, p->Wpi) * sizeof(double) * (1 + (p->bursts ? p->bursts->nB : 0));
    size_t chunk_mb = 384;
    if (const char* e = getenv("NMB200_CHUNK_MB")) chunk_mb = (size_t)std::max(1, atoi(e));  // tuning knob (profiling only)
    p->chunk = (int)std::max<size_t>(1, std::min<size_t>(64, (chunk_mb << 20) / per_window));
    std::vector<long long> yoff(p->chunk);
    for (int k = 0; k < p->chunk; ++k) yoff[k] = (long long)k * p->C * p->Wpi;
    if (p->d_yoff.upload(yoff, p->stream)) return -1;
    if (p->notch && p->d_y.ensure((size_t)p->chunk * p->C * p->Wpi * sizeof(double))) return -1;
    for (size_t i = 0; i < std::min<size_t>(2, p->prefilters.size()); ++i)
        if (p->d_pre[i].ensure((size_t)p->chunk * p->C * p->Wpi * sizeof(double) + 16)) return -1;
    if (p->has_resampler) {
        for (int k = 0; k < p->chunk; ++k) yoff[k] = (long long)k * p->C * p->Wp;
        if (p->d_roff.upload(yoff, p->stream)) return -1;
        if (p->d_rs.ensure((size_t)p->chunk * p->C * p->Wp * sizeof(double))) return -1;
        if (p->rs_bank && nm_allow_fir_smem<NmEpiStoreDecim>(*p->rs_bank, 0, p)) return -1;
    }
    for (auto& b : p->prefilters)
        if (nm_allow_fir_smem<NmEpiStore>(*b, 0, p)) return -1;
    if (p->bursts && p->bursts->alloc_chunk(p->chunk, p->Wp)) return -1;
    if (p->rawnorm && p->rawnorm->alloc_chunk(p->chunk, p->Wp)) return -1;
    if (p->rawnorm && p->rawnorm->need_median() && nm_allow_smem(nm_burst_thr_kernel, BurstsFam::thr_smem(), p)) return -1;

    // bulk-copy staged notch rows: needs the specialised single-filter reflect plan and room for the stage
    p->notch_tma = false;
    if (p->notch && nm_convx_pick<NmEpiStoreScan>(*p->notch)) {
        const char* env = getenv("NMB200_NOTCH_TMA");
        const bool want = env ? atoi(env) != 0 : true;
        if (want && nm_notchx_smem(p->notch->P, p->Win) <= (size_t)p->smem_max) {
            if (nm_allow_smem(nm_notchx_pick(p->notch->P), nm_notchx_smem(p->notch->P, p->Win), p)) return -1;
            p->notch_tma = true;
        }
    }
    // opt in to large dynamic shared memory once
    if (p->notch && (nm_allow_fir_smem<NmEpiStore>(*p->notch, 0, p) || nm_allow_fir_smem<NmEpiStoreScan>(*p->notch, 0, p) ||
                     nm_allow_fir_smem<NmEpiStoreScan>(*p->notch, 0, p, 1) || nm_allow_fir_smem<NmEpiStoreScan>(*p->notch, 0, p, 2)))
        return -1;
    if (p->bandpower && (nm_allow_fir_smem<NmEpiBandpower>(p->bandpower->bank, p->bandpower->epi_smem(), p) ||
                         nm_allow_fir_smem<NmEpiBandpower>(p->bandpower->bank, p->bandpower->epi_smem(), p, 1) ||
                         nm_allow_fir_smem<NmEpiBandpower>(p->bandpower->bank, p->bandpower->epi_smem(), p, 2)))
        return -1;
    size_t spec_max = 0;
    for (auto& f : p->spectral)
        if (!f->fast) spec_max = std::max(spec_max, f->smem());
    if (spec_max && nm_allow_smem(nm_spec_kernel, spec_max, p)) return -1;
    for (auto& f : p->spectral)
        if (f->fast && nm_allow_smem(f->kernel(), f->smem(), p)) return -1;
    if (p->bursts && p->bursts->allow_smem(p)) return -1;
    if (p->sharpwave && p->sharpwave->allow_smem(p)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    p->finalized = true;
    return 0;
}

extern "C" int nm_reset_state(nm_pipeline* p) {
    NM_P_CHECK(p);
    cudaSetDevice(p->device);
    if (p->bursts) p->bursts->reset(p->stream);
    if (p->norm) p->norm->reset();
    if (p->rawnorm) p->rawnorm->reset(p->stream);
    return 0;
}

// ------------------------------------------------------------------------------- data path
static NmPrepArgs nm_prep_args(nm_pipeline* p) {
    NmPrepArgs a;
    a.raw = p->d_raw.p;
    a.raw_is_f64 = p->raw_f64 ? 1 : 0;
    a.raw_pitch = p->raw_pitch;
    a.C_all = p->C_all;
    a.T = p->T;
    a.C = p->C;
    a.pick = p->d_pick.as<int>();
    a.G = p->G;
    a.group_of = p->d_group_of.as<int>();
    a.gcoef = p->d_gcoef.as<double>();
    a.sp_ptr = p->d_sp_ptr.as<int>();
    a.sp_col = p->d_sp_col.as<int>();
    a.sp_val = p->d_sp_val.as<double>();
    a.xr = p->d_xr.as<double>();
    a.xr_pitch = p->xr_pitch;
    a.nanblk = p->d_nanblk.as<unsigned char>();
    a.nanblk_pitch = p->nanblk_pitch;
    a.gsum_ext = nullptr;
    a.gsum_pitch = 0;
    // fused pipelines keep only the group sums (the re-reference is folded into the window kernel's load)
    const bool fused = p->fused && !p->force_xr;
    a.write_xr = fused ? 0 : 1;
    a.gsum_out = (fused && p->G > 0) ? p->d_gsum.as<double>() : nullptr;
    if (a.gsum_out) a.gsum_pitch = p->gsum_pitch;
    a.t0 = 0;
    a.t1 = p->T;
    return a;
}

#define NM_UPLOAD_SLICES 8
#define NM_UPLOAD_MIN_PIPELINED (1 << 16)  // samples; shorter recordings are copied and re-referenced in one go

// re-reference every uploaded time slice that holds samples below `upto` and has not been processed yet
static void nm_upload_join(nm_pipeline* p) {
    if (p->up_thread.joinable()) p->up_thread.join();
    p->up_active = false;
}
// block until the upload thread has enqueued slice k (its event is recorded); false if the thread failed
static bool nm_upload_wait_enqueued(nm_pipeline* p, int k) {
    if (!p->up_active) return true;
    std::unique_lock<std::mutex> lk(p->up_mu);
    p->up_cv.wait(lk, [&] { return p->up_enqueued > k || p->up_failed; });
    return !p->up_failed;
}

static int nm_ensure_prepped(nm_pipeline* p, long long upto) {
    while (p->slices_prepped < p->n_slices && (long long)p->slices_prepped * p->slice_len < upto) {
        const int k = p->slices_prepped;
        NmPrepArgs a = nm_prep_args(p);
        if (p->resident_uses_gsum) {  // channel-sharded upload: the slice's group sums must have been all-reduced
            NM_CHECK(k < p->slices_reduced, "slice %d of the sharded upload has not been reduced yet (nm_upload_slice_reduced)", k);
            NM_CUDA_CHECK(cudaStreamWaitEvent(p->stream, p->red_ev[k], 0));
            a.gsum_ext = p->d_gsum.as<double>();
            a.gsum_pitch = p->gsum_pitch;
            a.gsum_out = nullptr;
        } else {
            NM_CHECK(nm_upload_wait_enqueued(p, k), "the deferred upload failed");
            NM_CUDA_CHECK(cudaStreamWaitEvent(p->stream, p->slice_ev[k], 0));
        }
        a.t0 = (long long)k * p->slice_len;
        a.t1 = std::min<long long>(p->T, a.t0 + p->slice_len);
        nm_launch_prep(p, a);
        p->slices_prepped++;
    }
    NM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// device buffers of a recording whose geometry (T, pitches) has just been set; a fused pipeline keeps the group sums instead of
// the re-referenced float64 copy.  The raw rows get 16 spare bytes: the fused kernel's bulk copies round sizes up to 16 bytes.
static int nm_stage_buffers(nm_pipeline* p, size_t esz) {
    if (p->d_raw.ensure((size_t)p->C_all * p->raw_pitch * esz + 16)) return -1;
    const bool fused = p->fused && !p->force_xr;
    if (!fused && p->d_xr.ensure((size_t)p->C * p->xr_pitch * sizeof(double) + 16)) return -1;
    if (p->d_nanblk.ensure((size_t)p->C_all * p->nanblk_pitch)) return -1;
    if (fused && p->G > 0) {
        p->gsum_pitch = (p->T + 1) & ~1LL;
        if (p->d_gsum.ensure((size_t)p->G * p->gsum_pitch * sizeof(double) + 16)) return -1;
    }
    return 0;
}

static int nm_stage_raw(nm_pipeline* p, const void* data, bool f64, long long n_samples, long long pitch) {
    const size_t esz = f64 ? 8 : 4;
    p->n_slices = 0;
    p->slices_prepped = 0;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->copy_stream));  // a pipelined upload / download may still use the buffers
    p->raw_f64 = f64;
    p->T = n_samples;
    p->raw_pitch = (n_samples + 3) & ~3LL;
    p->xr_pitch = (n_samples + 1) & ~1LL;
    p->nanblk_pitch = (n_samples + 31) / 32;
    if (nm_stage_buffers(p, esz)) return -1;
    if (pitch == p->raw_pitch) {
        NM_CUDA_CHECK(cudaMemcpyAsync(p->d_raw.p, data, (size_t)p->C_all * pitch * esz, cudaMemcpyHostToDevice, p->stream));
    } else {
        for (int r = 0; r < p->C_all; ++r)
            NM_CUDA_CHECK(cudaMemcpyAsync((char*)p->d_raw.p + (size_t)r * p->raw_pitch * esz, (const char*)data + (size_t)r * pitch * esz,
                                          (size_t)n_samples * esz, cudaMemcpyHostToDevice, p->stream));
    }
    return 0;
}

// geometry + asynchronous H2D of the recording in `n_slices` time slices on the copy stream (slice_ev[k] = slice k has landed)
static int nm_stage_slices(nm_pipeline* p, const void* data, bool f64, long long n_samples, long long pitch, int n_slices, bool deferred = false) {
    const size_t esz = f64 ? 8 : 4;
    p->raw_f64 = f64;
    p->T = n_samples;
    p->raw_pitch = (n_samples + 3) & ~3LL;
    p->xr_pitch = (n_samples + 1) & ~1LL;
    p->nanblk_pitch = (n_samples + 31) / 32;
    if (nm_stage_buffers(p, esz)) return -1;
    p->slice_len = ((n_samples + n_slices - 1) / n_slices + 255) & ~255LL;
    p->n_slices = (int)((n_samples + p->slice_len - 1) / p->slice_len);
    p->slices_prepped = 0;
    p->slices_reduced = 0;
    while ((int)p->slice_ev.size() < p->n_slices) {
        cudaEvent_t e;
        NM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->slice_ev.push_back(e);
        NM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->red_ev.push_back(e);
    }
    // kernels of the previous run may still read the buffers that are about to be overwritten
    NM_CUDA_CHECK(cudaEventRecord(p->ev_sync, p->stream));
    NM_CUDA_CHECK(cudaStreamWaitEvent(p->copy_stream, p->ev_sync, 0));
    if (deferred) {
        // two page-locked staging buffers of one slice each; the thread below owns them and the copy stream's H2D work until it ends
        const size_t need = (size_t)p->C_all * p->slice_len * esz;
        if (need > p->up_stage_bytes) {
            p->up_stage_bytes = 0;
            bool got = true;
            for (int b = 0; b < 2; ++b) {
                if (p->up_stage[b]) cudaFreeHost(p->up_stage[b]);
                p->up_stage[b] = nullptr;
                got = got && cudaMallocHost(&p->up_stage[b], need) == cudaSuccess;
                if (!p->up_free[b]) NM_CUDA_CHECK(cudaEventCreateWithFlags(&p->up_free[b], cudaEventDisableTiming));
            }
            if (got) p->up_stage_bytes = need;
            else {  // no page-locked memory for the staging buffers: the driver's own staged copies below
                cudaGetLastError();
                for (int b = 0; b < 2; ++b) {
                    if (p->up_stage[b]) cudaFreeHost(p->up_stage[b]);
                    p->up_stage[b] = nullptr;
                }
                deferred = false;
            }
        }
    }
    if (deferred) {
        p->up_enqueued = 0;
        p->up_failed = false;
        p->up_active = true;
        auto body = [p, data, esz, n_samples, pitch]() {
            cudaSetDevice(p->device);
            bool ok = true;
            for (int k = 0; k < p->n_slices && ok; ++k) {
                const int b = k & 1;
                const long long t0 = (long long)k * p->slice_len, len = std::min<long long>(p->slice_len, n_samples - t0);
                if (k >= 2) ok = ok && cudaEventSynchronize(p->up_free[b]) == cudaSuccess;  // slice k - 2 has left this buffer
                char* st = (char*)p->up_stage[b];
                // (one thread moves ~6 GB/s out of pageable memory: the rows of a slice are dealt to a few helpers)
                auto pack = [&](int r0, int r1) {
                    for (int r = r0; r < r1; ++r)
                        memcpy(st + (size_t)r * len * esz, (const char*)data + ((size_t)r * pitch + t0) * esz, (size_t)len * esz);
                };
                const int nh = (size_t)p->C_all * len * esz >= ((size_t)4 << 20) ? ((int)std::thread::hardware_concurrency() >= 4 ? 4 : 2) : 1;
                if (nh > 1) {
                    std::vector<std::thread> helpers;
                    for (int h = 1; h < nh; ++h) helpers.emplace_back(pack, (int)((long long)p->C_all * h / nh), (int)((long long)p->C_all * (h + 1) / nh));
                    pack(0, p->C_all / nh);
                    for (auto& t : helpers) t.join();
                } else {
                    pack(0, p->C_all);
                }
                ok = ok && cudaMemcpy2DAsync((char*)p->d_raw.p + (size_t)t0 * esz, (size_t)p->raw_pitch * esz, st, (size_t)len * esz,
                                             (size_t)len * esz, (size_t)p->C_all, cudaMemcpyHostToDevice, p->copy_stream) == cudaSuccess;
                ok = ok && cudaEventRecord(p->up_free[b], p->copy_stream) == cudaSuccess;
                ok = ok && cudaEventRecord(p->slice_ev[k], p->copy_stream) == cudaSuccess;
                {
                    std::lock_guard<std::mutex> lk(p->up_mu);
                    if (ok) p->up_enqueued = k + 1;
                    else p->up_failed = true;
                }
                p->up_cv.notify_all();
            }
        };
#ifdef NM_EMULATE
        body();  // (the test build copies synchronously)
#else
        p->up_thread = std::thread(body);
#endif
        return 0;
    }
    for (int k = 0; k < p->n_slices; ++k) {
        const long long t0 = (long long)k * p->slice_len, len = std::min<long long>(p->slice_len, n_samples - t0);
        NM_CUDA_CHECK(cudaMemcpy2DAsync((char*)p->d_raw.p + (size_t)t0 * esz, (size_t)p->raw_pitch * esz, (const char*)data + (size_t)t0 * esz,
                                        (size_t)pitch * esz, (size_t)len * esz, (size_t)p->C_all, cudaMemcpyHostToDevice, p->copy_stream));
        NM_CUDA_CHECK(cudaEventRecord(p->slice_ev[k], p->copy_stream));
    }
    return 0;
}

static int nm_upload_impl(nm_pipeline* p, const void* data, bool f64, long long n_samples, long long pitch, bool deferred = false) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized, "call nm_finalize first");
    NM_CHECK(data && n_samples >= p->Win && pitch >= n_samples, "bad recording geometry (n_samples %lld, pitch %lld, W %d)", n_samples,
             pitch, p->Win);
    cudaSetDevice(p->device);
    nm_upload_join(p);     // (a deferred upload that was never consumed)
    nm_stream_release(p);  // a streaming session (captured graphs, one-window geometry) does not survive a batched upload
    p->n_slices = 0;
    p->slices_prepped = 0;
    if (n_samples >= NM_UPLOAD_MIN_PIPELINED) {
        // pipelined upload: time slices on the copy stream, each followed by an event; nm_ensure_prepped() makes the compute
        // stream wait for (and re-reference) a slice only when a chunk of windows first needs it.  Deferred: only for pageable
        // memory (page-locked buffers are copied asynchronously by the copy engine as they are)
#ifndef NM_EMULATE
        if (deferred) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, data) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) deferred = false;
            cudaGetLastError();
        }
#endif
        if (nm_stage_slices(p, data, f64, n_samples, pitch, NM_UPLOAD_SLICES, deferred)) return -1;
    } else {
        if (nm_stage_raw(p, data, f64, n_samples, pitch)) return -1;
        nm_launch_prep(p, nm_prep_args(p));
        NM_CUDA_CHECK(cudaGetLastError());
    }
    p->have_data = true;
    p->upload_pending = false;
    p->resident_uses_gsum = false;
    return 0;
}

// re-run the window-independent preprocessing on the recording that is already resident on the device
extern "C" int nm_prepare_resident(nm_pipeline* p) {
    NM_P_CHECK(p);
    NM_CHECK(p->have_data && !p->upload_pending, "no resident recording");
    cudaSetDevice(p->device);
    if (nm_ensure_prepped(p, p->T)) return -1;  // (a pipelined upload may still be in flight)
    NmPrepArgs a = nm_prep_args(p);
    if (p->resident_uses_gsum) {  // channel-sharded recording: keep using the all-reduced group sums
        a.gsum_ext = p->d_gsum.as<double>();
        a.gsum_pitch = p->gsum_pitch;
        a.gsum_out = nullptr;
    }
    nm_launch_prep(p, a);
    NM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// pageable recordings are staged by a host thread (deferred upload) unless NMB200_DEFERRED_UPLOAD=0
static bool nm_deferred_upload_enabled() {
    const char* env = getenv("NMB200_DEFERRED_UPLOAD");
    return env ? atoi(env) != 0 : true;
}
extern "C" int nm_upload_f32(nm_pipeline* p, const float* data, long long n_samples, long long pitch) {
    return nm_upload_impl(p, data, false, n_samples, pitch, nm_deferred_upload_enabled());
}
extern "C" int nm_upload_f64(nm_pipeline* p, const double* data, long long n_samples, long long pitch) {
    return nm_upload_impl(p, data, true, n_samples, pitch, nm_deferred_upload_enabled());
}


// PreprocessingFilter stages of one batch of windows; on return `rows` describes the last stage's output buffer
static void nm_run_prefilters(nm_pipeline* p, NmRows& rows) {
    for (size_t i = 0; i < p->prefilters.size(); ++i) {
        double* dst = p->d_pre[i & 1].as<double>();
        NmEpiStore epi{dst, (long long)p->Wpi, 1};
        p->prof_begin();
        nm_launch_fir(p, *p->prefilters[i], rows, epi, p->stream, 0);
        p->prof_end(NM_PROF_NOTCH);
        p->launches++;
        rows.base = dst;
        rows.ch_stride = p->Wpi;
        rows.off = p->d_yoff.as<long long>();
    }
}

// raw_resampling of one batch of windows: rows of Win samples -> d_rs rows of W samples (nm_resample.cuh)
static void nm_run_resampler(nm_pipeline* p, NmRows& rows) {
    NmResampleArgs a;
    a.in = rows;
    a.rt = p->d_rt.as<double>();
    a.n_out = p->W;
    a.n_out_pitch = p->rs_pitch;
    a.out = p->d_rs.as<double>();
    a.out_pitch = p->Wp;
    const long long n_rows = (long long)rows.n_windows * rows.n_ch;
    const dim3 grid((unsigned)((n_rows + NM_RS_BM - 1) / NM_RS_BM), (unsigned)((p->W + NM_RS_BN - 1) / NM_RS_BN));
    p->prof_begin();
    if (p->rs_bank) {  // integer down-sampling: ideal low-pass by FFT convolution, every D-th sample stored
        NmEpiStoreDecim epi{p->d_rs.as<double>(), (long long)p->Wp, p->rs_decim};
        nm_launch_fir(p, *p->rs_bank, rows, epi, p->stream, 0);
    } else {
        NM_LAUNCH(nm_resample_kernel, grid, dim3(NM_RS_THREADS), nm_resample_smem_bytes(), p->stream, a);
    }
    p->prof_end(NM_PROF_RESAMPLE);
    p->launches++;
    rows.base = p->d_rs.as<double>();
    rows.ch_stride = p->Wp;
    rows.off = p->d_roff.as<long long>();
    rows.W = p->W;
}

static int nm_run_chunk(nm_pipeline* p, int w0, int n) {
    NmRows rows;
    rows.base = p->d_xr.as<double>();
    rows.ch_stride = p->xr_pitch;
    rows.off = p->d_starts.as<long long>() + w0;
    rows.n_windows = n;
    rows.n_ch = p->C;
    rows.W = p->Win;

    auto out_for = [&](const DevBuf& colmap, int per_ch) {
        NmOut o;
        o.out = p->d_out.as<double>();
        o.row0 = w0;
        o.F = p->F;
        o.colmap = colmap.as<int>();
        o.per_ch = per_ch;
        return o;
    };
    nm_run_prefilters(p, rows);
    bool scan_done = false;
    if (p->fused) {
        // ---- one kernel: folded re-reference, notch, scan, in-kernel DFT families, band-pass bank (nm_fused.cuh)
        bool y_needed = p->sharpwave || p->bursts || (p->bandpower && !p->fused_bp) || !(p->has_scan || !p->fused_spec.empty() || p->fused_bp);
        for (size_t i = 0; i < p->spectral.size(); ++i) y_needed = y_needed || !nm_spec_in_fused(p, i);
        const FirBank& nb = *p->notch;
        NmFusedArgs a;
        a.raw = p->d_raw.p;
        a.raw_pitch = p->raw_pitch;
        a.pick = p->d_pick.as<int>();
        a.dcoef = p->d_dfold.as<double>();
        a.gcoef = p->d_gfold.as<double>();
        a.gsum = p->G > 0 ? p->d_gsum.as<double>() : nullptr;
        a.start = p->d_starts.as<long long>() + w0;
        a.n_windows = n; a.n_ch = p->C; a.W = p->W; a.E = nb.E;
        a.n_items = n * ((p->C + 1) / 2);
        a.tw = nb.fft.d_tw.as<cx<double>>();
        a.hx_notch = nb.d_hx.as<double>();
        a.hx_bank = nullptr;
        a.nF = 0;
        a.scan.y = y_needed ? p->d_y.as<double>() : nullptr;
        a.scan.Wp = p->Wp;
        a.scan.want_scan = p->has_scan ? 1 : 0;
        a.scan.want_hjorth = p->scan_h; a.scan.want_raw = p->scan_r; a.scan.want_ll = p->scan_l;
        a.scan.out = p->has_scan ? out_for(p->d_scan_colmap, 5) : NmOut{nullptr, 0, 0, nullptr, 0};
        a.bp.seglen = nullptr;
        a.bp.want_act = a.bp.want_mob = a.bp.want_comp = a.bp.log_act = 0;
        a.bp.out = NmOut{nullptr, 0, 0, nullptr, 0};
        for (int b = 0; b < NM_BP_MAX_INLINE; ++b) a.bp.seglen_k[b] = 1;
        if (p->fused_bp) {
            BandpowerFam& f = *p->bandpower;
            a.hx_bank = f.bank.d_hx.as<double>();
            a.nF = f.bank.nF;
            a.bp.seglen = f.d_seglen.as<int>();
            for (int b = 0; b < NM_BP_MAX_INLINE; ++b) a.bp.seglen_k[b] = b < (int)f.h_seglen.size() ? f.h_seglen[b] : 1;
            a.bp.want_act = f.act; a.bp.want_mob = f.mob; a.bp.want_comp = f.comp; a.bp.log_act = f.logt;
            a.bp.out = out_for(f.d_colmap, f.bank.nF * 3);
        }
        a.n_spec = (int)p->fused_spec.size();
        a.spec = p->d_fused_spec.as<NmSpecArgs>();
        a.row0 = w0;
        auto kern = p->front ? nm_front_pick(nb.P, p->fused_sx, p->raw_f64) : nm_fused_pick(nb.P, p->fused_sx, p->raw_f64);
        const size_t sm = p->front ? nm_front_smem(nb.P, p->W, p->raw_f64) : nm_fused_smem(nb.P);
        const int threads = nb.P / 16;
        p->prof_begin();
        NM_LAUNCH(kern, dim3(nm_resident_grid(p, kern, threads, sm, a.n_items)), dim3(threads), sm, p->stream, a);
        p->prof_end(NM_PROF_FUSED);
        p->launches++;
        scan_done = true;
        rows.base = p->d_y.as<double>();
        rows.ch_stride = p->Wp;
        rows.off = p->d_yoff.as<long long>();
    } else if (p->notch) {
        const bool y_needed = !p->spectral.empty() || p->bandpower || p->sharpwave || p->bursts || p->rawnorm || p->has_resampler;
        // the scan features are taken from the NORMALISED / RESAMPLED rows otherwise
        const bool fuse_scan = p->has_scan && !p->rawnorm && !p->has_resampler;
        NmOut so;
        if (fuse_scan) so = out_for(p->d_scan_colmap, 5);
        p->prof_begin();
        scan_done = nm_launch_notch(p, rows, y_needed || !p->has_scan ? p->d_y.as<double>() : nullptr, fuse_scan ? &so : nullptr, p->stream);
        p->prof_end(NM_PROF_NOTCH);
        p->launches++;
        rows.base = p->d_y.as<double>();
        rows.ch_stride = p->Wpi;
        rows.off = p->d_yoff.as<long long>();
    }
    if (p->has_resampler) nm_run_resampler(p, rows);
    if (p->rawnorm && p->rawnorm->run(p, rows)) return -1;
    if (p->has_scan && !scan_done) {
        NmScanArgs a;
        a.in = rows;
        a.want_hjorth = p->scan_h; a.want_raw = p->scan_r; a.want_ll = p->scan_l;
        a.out = out_for(p->d_scan_colmap, 5);
        const int wpc = NM_ROW_THREADS / 32;
        const long long n_rows = (long long)n * p->C;
        const int grid = (int)std::max<long long>(1, std::min<long long>((n_rows + wpc - 1) / wpc, (long long)p->n_sm * 16));
        p->prof_begin();
        NM_LAUNCH(nm_scan_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, a);
        p->prof_end(NM_PROF_SCAN);
        p->launches++;
    }
    // ---- fork: the families below only read `rows` and write disjoint columns of `out`
    cudaStream_t const main_stream = p->stream;
    const int n_branches = ((!p->spectral.empty() || p->bandpower) ? 1 : 0) + (p->sharpwave ? 1 : 0) + (p->bursts ? 1 : 0);
    const bool fork = !p->profiling && n_branches >= 2;
    const bool live = !nm_gs_updating();  // (a replayed graph already holds these dependencies)
    if (fork && live) NM_CUDA_CHECK(cudaEventRecord(p->ev_fork, main_stream));
    auto branch_begin = [&](int b) {
        if (!fork) return;
        if (live) cudaStreamWaitEvent(p->side[b], p->ev_fork, 0);
        p->stream = p->side[b];
    };
    auto branch_end = [&](int b) {
        if (!fork) return;
        if (live) cudaEventRecord(p->ev_join[b], p->side[b]);
        p->stream = main_stream;
        if (live) cudaStreamWaitEvent(main_stream, p->ev_join[b], 0);
    };
    branch_begin(0);
    for (size_t fi = 0; fi < p->spectral.size(); ++fi) {
        auto& f = p->spectral[fi];
        if (p->fused && nm_spec_in_fused(p, fi)) continue;
        NmSpecArgs a = nm_spec_args(p, *f, rows, out_for(f->d_colmap, f->per_ch), n);
        const size_t sm = f->smem();
        p->prof_begin();
        if (f->fast) {
            auto k = f->kernel();
            NM_LAUNCH(k, dim3(nm_resident_grid(p, k, f->threads(), sm, a.n_items)), dim3(f->threads()), sm, p->stream, a);
        } else {
            NM_LAUNCH(nm_spec_kernel, dim3(p->grid_for(sm, a.n_items, NM_FFT_THREADS)), dim3(NM_FFT_THREADS), sm, p->stream, a);
        }
        p->prof_end(NM_PROF_SPEC);
        p->launches++;
    }
    if (p->bandpower && !(p->fused && p->fused_bp)) {
        BandpowerFam& f = *p->bandpower;
        NmEpiBandpower epi;
        epi.seglen = f.d_seglen.as<int>();
        for (int b = 0; b < NM_BP_MAX_INLINE; ++b) epi.seglen_k[b] = b < (int)f.h_seglen.size() ? f.h_seglen[b] : 1;
        epi.want_act = f.act; epi.want_mob = f.mob; epi.want_comp = f.comp; epi.log_act = f.logt;
        epi.out = out_for(f.d_colmap, f.bank.nF * 3);
        p->prof_begin();
        nm_launch_fir(p, f.bank, rows, epi, p->stream, f.epi_smem(), (f.mob || f.comp) ? 0 : p->precision);
        p->prof_end(NM_PROF_BANDPOWER);
        p->launches++;
    }
    branch_end(0);
    int rc = 0;
    branch_begin(1);
    if (p->sharpwave) rc |= p->sharpwave->run(p, rows, w0);
    branch_end(1);
    branch_begin(2);
    if (p->bursts) rc |= p->bursts->run(p, rows, w0);
    branch_end(2);
    if (rc) return -1;
    NM_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// argument blocks of the in-kernel DFT families of the fused / front kernel (they hold the d_out pointer: refreshed whenever d_out
// may have been re-allocated)
static int nm_fused_spec_upload(nm_pipeline* p) {
    if (!p->fused || p->fused_spec.empty()) return 0;
    p->h_fused_spec.clear();
    NmRows none{};
    for (int k : p->fused_spec) {
        const SpectralFam& f = *p->spectral[k];
        NmOut o;
        o.out = p->d_out.as<double>();
        o.row0 = 0;
        o.F = p->F;
        o.colmap = f.d_colmap.as<int>();
        o.per_ch = f.per_ch;
        p->h_fused_spec.push_back(nm_spec_args(p, f, none, o, 0));
    }
    if (p->d_fused_spec.upload(p->h_fused_spec, p->stream)) return -1;
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

// per-window NaN-channel flags from the block map + NaN re-insertion by column list for windows [w0, w0 + n) of the current run
static void nm_nan_fill(nm_pipeline* p, int w0, int n) {
    NmNanArgs na;
    na.p = nm_prep_args(p);
    na.start = p->d_starts.as<long long>() + w0;
    na.n_windows = n;
    na.W = p->Win;
    na.flags = p->d_nanflags.as<unsigned char>() + (size_t)w0 * p->C_all;
    const long long tot = (long long)n * p->C_all;
    const unsigned grid = (unsigned)((tot + NM_ROW_THREADS - 1) / NM_ROW_THREADS);
    NmNanFillArgs nf;
    nf.flags = na.flags;
    nf.n_windows = n;
    nf.C_all = p->C_all;
    nf.col_ptr = p->d_nan_ptr.as<int>();
    nf.cols = p->d_nan_cols.as<int>();
    nf.out = p->d_out.as<double>();
    nf.row0 = w0;
    nf.F = p->F;
    NM_LAUNCH(nm_nanfix_kernel, dim3(grid), dim3(NM_ROW_THREADS), 0, p->stream, na, nf);
    p->launches += 1;
}

extern "C" int nm_run_windows(nm_pipeline* p, const long long* starts, int n_windows, double* out_host) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized && p->have_data, "finalize the pipeline and upload a recording first");
    NM_CHECK(starts && n_windows > 0, "no windows given");
    nm_stream_release(p);
    for (int k = 0; k < n_windows; ++k)
        NM_CHECK(starts[k] >= 0 && starts[k] + p->Win <= p->T, "window %d [%lld, %lld) outside the recording (%lld samples)", k, starts[k],
                 starts[k] + p->Win, p->T);
    cudaSetDevice(p->device);
    if (p->d_starts.upload(starts, (size_t)n_windows, p->stream)) return -1;
    if (p->d_out.ensure((size_t)n_windows * p->F * sizeof(double))) return -1;
    if (nm_fused_spec_upload(p)) return -1;
    p->out_rows = n_windows;
    NM_CUDA_CHECK(cudaMemsetAsync(p->d_out.p, 0, (size_t)n_windows * p->F * sizeof(double), p->stream));
    // without the (sequential) normaliser a chunk's rows are final when its kernels end: ship them chunk by chunk
    const bool norm_per_chunk = p->norm && p->norm->per_chunk_ok();
    const bool per_chunk = (!p->norm || norm_per_chunk) && out_host != nullptr;
    bool stage_out = false;
#ifndef NM_EMULATE
    if (per_chunk && nm_deferred_upload_enabled()) {
        cudaPointerAttributes at;
        stage_out = cudaPointerGetAttributes(&at, out_host) == cudaSuccess && at.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        const size_t need = (size_t)n_windows * p->F * sizeof(double);
        if (stage_out && need > p->out_stage_bytes) {
            if (p->out_stage) cudaFreeHost(p->out_stage);
            p->out_stage = nullptr;
            p->out_stage_bytes = 0;
            if (cudaMallocHost(&p->out_stage, need) == cudaSuccess) p->out_stage_bytes = need;
            else { stage_out = false; cudaGetLastError(); }
        }
    }
#endif
    std::vector<std::pair<int, int>> staged;  // (first window, windows) of the chunks copied into out_stage
    if (p->has_nan_cols && p->d_nanflags.ensure((size_t)n_windows * p->C_all)) return -1;
    auto nan_fill = [&](int w0, int n) { nm_nan_fill(p, w0, n); };
    if (p->bursts && p->bursts->prepare(p, n_windows)) return -1;
    int n_ev = 0;
    for (int w0 = 0; w0 < n_windows; w0 += p->chunk) {
        const int n = std::min(p->chunk, n_windows - w0);
        long long upto = 0;
        for (int k = 0; k < n; ++k) upto = std::max(upto, starts[w0 + k] + p->Win);
        if (nm_ensure_prepped(p, upto)) return -1;
        if (nm_run_chunk(p, w0, n)) return -1;
        if (norm_per_chunk && p->norm->run(p, n, w0, n_windows)) return -1;
        if (per_chunk) {
            if (p->has_nan_cols) nan_fill(w0, n);
            if (out_host) {  // ship the finished rows while the next chunk computes
                if ((int)p->chunk_ev.size() <= n_ev) {
                    cudaEvent_t e;
                    NM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    p->chunk_ev.push_back(e);
                }
                NM_CUDA_CHECK(cudaEventRecord(p->chunk_ev[n_ev], p->stream));
                NM_CUDA_CHECK(cudaStreamWaitEvent(p->copy_stream, p->chunk_ev[n_ev], 0));
                const size_t hp = (size_t)(p->out_pitch ? p->out_pitch : p->F);
                if (stage_out) {
                    NM_CUDA_CHECK(cudaMemcpyAsync((double*)p->out_stage + (size_t)w0 * p->F, p->d_out.as<double>() + (size_t)w0 * p->F,
                                                  (size_t)n * p->F * sizeof(double), cudaMemcpyDeviceToHost, p->copy_stream));
                    if ((int)p->d2h_ev.size() <= n_ev) {
                        cudaEvent_t e;
                        NM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        p->d2h_ev.push_back(e);
                    }
                    NM_CUDA_CHECK(cudaEventRecord(p->d2h_ev[n_ev], p->copy_stream));
                    staged.emplace_back(w0, n);
                } else {
                    NM_CUDA_CHECK(cudaMemcpy2DAsync(out_host + (size_t)w0 * hp, hp * sizeof(double), p->d_out.as<double>() + (size_t)w0 * p->F,
                                                    (size_t)p->F * sizeof(double), (size_t)p->F * sizeof(double), (size_t)n,
                                                    cudaMemcpyDeviceToHost, p->copy_stream));
                }
                ++n_ev;
            }
        }
    }
    if (nm_ensure_prepped(p, p->T)) return -1;  // leave no slice event un-consumed (nm_prepare_resident, NaN maps)
    nm_upload_join(p);                           // (every slice of a deferred upload has been enqueued by now)
    if (!per_chunk) {
        if (p->norm && !norm_per_chunk && p->norm->run(p, n_windows)) return -1;
        if (p->has_nan_cols) nan_fill(0, n_windows);
    }
    NM_CUDA_CHECK(cudaGetLastError());
    if (out_host) {
        if (!per_chunk) return nm_download(p, out_host, n_windows);
        // rows of the staged chunks -> the caller's (pageable, possibly pitched) matrix, chunk by chunk as their copies land
        const size_t hp = (size_t)(p->out_pitch ? p->out_pitch : p->F);
        for (size_t i = 0; i < staged.size(); ++i) {
            NM_CUDA_CHECK(cudaEventSynchronize(p->d2h_ev[i]));
            const int w0 = staged[i].first, n = staged[i].second;
            const double* src = (const double*)p->out_stage + (size_t)w0 * p->F;
            if (hp == (size_t)p->F) memcpy(out_host + (size_t)w0 * hp, src, (size_t)n * p->F * sizeof(double));
            else
                for (int r = 0; r < n; ++r) memcpy(out_host + (size_t)(w0 + r) * hp, src + (size_t)r * p->F, (size_t)p->F * sizeof(double));
        }
        NM_CUDA_CHECK(cudaStreamSynchronize(p->copy_stream));
        NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    }
    return 0;
}

extern "C" int nm_download(nm_pipeline* p, double* out_host, int n_windows) {
    NM_P_CHECK(p);
    NM_CHECK(out_host && n_windows > 0 && n_windows <= p->out_rows, "nothing to download");
    cudaSetDevice(p->device);
    const size_t hp = (size_t)(p->out_pitch ? p->out_pitch : p->F);
    NM_CUDA_CHECK(cudaMemcpy2DAsync(out_host, hp * sizeof(double), p->d_out.p, (size_t)p->F * sizeof(double), (size_t)p->F * sizeof(double),
                                    (size_t)n_windows, cudaMemcpyDeviceToHost, p->stream));
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int nm_set_output_pitch(nm_pipeline* p, long long pitch_elems) {
    NM_P_CHECK(p);
    NM_CHECK(pitch_elems == 0 || pitch_elems >= p->F, "output pitch %lld is smaller than the %d features of a row", pitch_elems, p->F);
    p->out_pitch = pitch_elems;
    return 0;
}

extern "C" int nm_host_register(void* ptr, long long bytes) {
    NM_CHECK(ptr && bytes > 0, "bad arguments");
    NM_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
    return 0;
}
extern "C" int nm_host_unregister(void* ptr) {
    NM_CUDA_CHECK(cudaHostUnregister(ptr));
    return 0;
}

extern "C" int nm_stream_open(nm_pipeline* p, int n_slots, int input_f32, int use_graph);
extern "C" int nm_stream_submit(nm_pipeline* p, int slot);
extern "C" int nm_stream_wait(nm_pipeline* p, int slot, const double** features);

// One window with caller-owned (pageable) buffers: the synchronous form of the streaming entry -- copy into a page-locked slot,
// submit, wait, copy the feature row out.  Producers that can write into the slot themselves use nm_stream_* directly.
extern "C" int nm_process_window(nm_pipeline* p, const double* window, double* out_features) {
    NM_P_CHECK(p);
    NM_CHECK(window && out_features, "NULL argument");
    NM_CHECK(p->finalized, "call nm_finalize first");
    if (!p->strm || p->strm->f32) {  // (batched entry points close the stream: its graphs hold their buffers' old addresses)
        if (nm_stream_open(p, 2, 0, -1)) return -1;
    }
    NmStream& st = *p->strm;
    const int slot = (int)(st.windows % st.n_slots);
    if (st.slots[slot].busy && nm_stream_wait(p, slot, nullptr)) return -1;
    memcpy(st.slots[slot].in_host, window, (size_t)p->C_all * p->Win * sizeof(double));
    if (nm_stream_submit(p, slot)) return -1;
    const double* f = nullptr;
    if (nm_stream_wait(p, slot, &f)) return -1;
    memcpy(out_features, f, (size_t)p->F * sizeof(double));
    return 0;
}

// Preprocessed rows of one window (nan_to_num -> pick -> re-reference -> notch), float64 (n_ch, W).
// Needed by user-defined (Python) features that run next to the GPU families, and by the stand-alone
// ReReferencer / DataPreprocessor classes.
extern "C" int nm_preprocess_window(nm_pipeline* p, const double* window, double* out_rows) {
    NM_P_CHECK(p);
    NM_CHECK(window && out_rows, "NULL argument");
    p->force_xr = true;  // this entry hands out the re-referenced / notched ROWS: materialise them (un-fused kernels)
    const int rc_up = nm_upload_impl(p, window, true, p->Win, p->Win);
    p->force_xr = false;
    if (rc_up) return -1;
    const long long zero = 0;
    if (p->d_starts.upload(&zero, 1, p->stream)) return -1;
    NmRows rows;
    rows.base = p->d_xr.as<double>();
    rows.ch_stride = p->xr_pitch;
    rows.off = p->d_starts.as<long long>();
    rows.n_windows = 1;
    rows.n_ch = p->C;
    rows.W = p->Win;
    nm_run_prefilters(p, rows);
    if (p->notch) {
        nm_launch_notch(p, rows, p->d_y.as<double>(), nullptr, p->stream);
        p->launches++;
        rows.base = p->d_y.as<double>();
        rows.ch_stride = p->Wpi;
    }
    if (p->has_resampler) nm_run_resampler(p, rows);
    if (p->rawnorm && p->rawnorm->run(p, rows)) return -1;
    NM_CUDA_CHECK(cudaGetLastError());
    const double* src = rows.base;
    const long long pitch = rows.ch_stride;
    for (int c = 0; c < p->C; ++c)
        NM_CUDA_CHECK(cudaMemcpyAsync(out_rows + (size_t)c * p->W, src + (size_t)c * pitch, (size_t)p->W * sizeof(double),
                                      cudaMemcpyDeviceToHost, p->stream));
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}

// ------------------------------------------------------------------------------- measurement
extern "C" int nm_timer_start(nm_pipeline* p) {
    NM_P_CHECK(p);
    cudaSetDevice(p->device);
    NM_CUDA_CHECK(cudaEventRecord(p->ev0, p->stream));
    return 0;
}
extern "C" int nm_timer_stop(nm_pipeline* p, double* elapsed_ms) {
    NM_P_CHECK(p);
    cudaSetDevice(p->device);
    NM_CUDA_CHECK(cudaEventRecord(p->ev1, p->stream));
    NM_CUDA_CHECK(cudaEventSynchronize(p->ev1));
    float ms = 0;
    NM_CUDA_CHECK(cudaEventElapsedTime(&ms, p->ev0, p->ev1));
    if (elapsed_ms) *elapsed_ms = ms;
    return 0;
}
extern "C" long long nm_kernel_launches(nm_pipeline* p) { return p ? p->launches : 0; }
extern "C" int nm_set_profiling(nm_pipeline* p, int enabled) {
    NM_P_CHECK(p);
    p->profiling = enabled != 0;
    for (int i = 0; i < NM_PROF_N; ++i) { p->prof_ms[i] = 0; p->prof_cnt[i] = 0; }
    return 0;
}
extern "C" int nm_get_profile(nm_pipeline* p, double* ms, long long* launches, int n) {
    NM_P_CHECK(p);
    for (int i = 0; i < n && i < NM_PROF_N; ++i) {
        if (ms) ms[i] = p->prof_ms[i];
        if (launches) launches[i] = p->prof_cnt[i];
    }
    return NM_PROF_N;
}
extern "C" int nm_chunk_windows(nm_pipeline* p) { return p ? p->chunk : 0; }
extern "C" int nm_burst_threshold_stats(nm_pipeline* p, long long* rebuilds, long long* direct_windows) {
    NM_P_CHECK(p);
    long long r = 0, d = 0;
    if (p->bursts && p->bursts->qstate.allocated()) {
        cudaSetDevice(p->device);
        NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
        std::vector<NmBurstQRow> rows((size_t)p->bursts->C * p->bursts->nB);
        NM_CUDA_CHECK(cudaMemcpy(rows.data(), p->bursts->qstate.current(), rows.size() * sizeof(NmBurstQRow), cudaMemcpyDeviceToHost));
        for (const auto& q : rows) { r += q.rebuilds; d += q.directs; }
    }
    if (rebuilds) *rebuilds = r;
    if (direct_windows) *direct_windows = d;
    return 0;
}
extern "C" int nm_set_burst_threshold_mode(nm_pipeline* p, int incremental) {
    NM_P_CHECK(p);
    NM_CHECK(p->finalized, "call nm_finalize first");
    if (p->bursts) {
        cudaSetDevice(p->device);
        NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
        p->bursts->incremental = incremental ? 1 : 0;
        p->bursts->reset(p->stream);
    }
    return 0;
}

template <class Epi>
static void nm_describe_fir(std::string& s, const char* family, const FirBank& b, size_t epi_bytes) {
    char line[256];
    if constexpr (Epi::kSplitOk) {
        if (b.nF > 1 && nm_split_banks_enabled() && nm_convx_pick_single<Epi>(b)) {
            snprintf(line, sizeof(line), "%s: nm_convx_kernel P=%d filters=%d (one launch per filter, single transform buffer) taps=%d threads=%d smem=%zu\n",
                     family, b.P, b.nF, b.L, b.threads(), b.smem_x(epi_bytes, 0, true));
            s += line;
            return;
        }
    }
    const bool x = nm_convx_pick<Epi>(b) != nullptr;
    snprintf(line, sizeof(line), "%s: %s P=%d filters=%d taps=%d threads=%d smem=%zu\n", family,
             x ? "nm_convx_kernel" : (b.pow2 ? "nm_conv_kernel" : "nm_fir_kernel"), b.P, b.nF, b.L, b.threads(), nm_fir_smem<Epi>(b, epi_bytes));
    s += line;
}

extern "C" int nm_describe_plan(nm_pipeline* p, char* buf, int n) {
    if (!p || !p->finalized) return 0;
    std::string s;
    char line[256];
    snprintf(line, sizeof(line), "window=%d channels=%d features=%d chunk=%d\n", p->W, p->C, p->F, p->chunk);
    s += line;
    if (p->has_resampler) {
        if (p->rs_bank)
            snprintf(line, sizeof(line), "resampler: nm_convx_kernel %d -> %d samples per row (P=%d ideal low-pass, every %d-th sample stored)\n",
                     p->Win, p->W, p->rs_bank->P, p->rs_decim);
        else
            snprintf(line, sizeof(line), "resampler: nm_resample_kernel %d -> %d samples per row (dense float64 operator)\n", p->Win, p->W);
        s += line;
    }
    for (auto& b : p->prefilters) nm_describe_fir<NmEpiStore>(s, "prefilter", *b, 0);
    if (p->fused) {
        snprintf(line, sizeof(line), "fused: %s P=%d threads=%d smem=%zu [re-reference + notch%s%s%s] dft_families=%d\n",
                 p->front ? "nm_front_kernel" : "nm_fused_kernel", p->notch->P, p->notch->P / 16,
                 p->front ? nm_front_smem(p->notch->P, p->W, p->raw_f64) : nm_fused_smem(p->notch->P), p->has_scan ? " + scan" : "",
                 p->fused_sx ? " + segment DFT" : "", p->fused_bp ? " + band-pass bank" : "", (int)p->fused_spec.size());
        s += line;
    } else if (p->notch) {
        if (nm_convx_pick<NmEpiStoreScan>(*p->notch) && p->notch_tma && !p->f32_linear()) {
            snprintf(line, sizeof(line), "%s: nm_notchx_kernel P=%d filters=1 taps=%d threads=%d smem=%zu (rows staged by cp.async.bulk + mbarrier)\n",
                     p->has_scan ? "notch+scan" : "notch", p->notch->P, p->notch->L, p->notch->threads(), nm_notchx_smem(p->notch->P, p->Win));
            s += line;
        } else if (nm_convx_pick<NmEpiStoreScan>(*p->notch)) nm_describe_fir<NmEpiStoreScan>(s, p->has_scan ? "notch+scan" : "notch", *p->notch, 0);
        else nm_describe_fir<NmEpiStore>(s, "notch", *p->notch, 0);
    }
    if (p->has_scan && !p->fused && !(p->notch && nm_convx_pick<NmEpiStoreScan>(*p->notch))) s += "scan: nm_scan_kernel\n";
    for (size_t fi = 0; fi < p->spectral.size(); ++fi) {
        auto& f = p->spectral[fi];
        if (p->fused && nm_spec_in_fused(p, fi)) continue;
        snprintf(line, sizeof(line), "spectral: %s nper=%d nseg=%d bins=%d threads=%d smem=%zu\n", f->fast ? "nm_specx_kernel" : "nm_spec_kernel",
                 f->cfg.nper, f->cfg.nseg, f->nk, f->threads(), f->smem());
        s += line;
    }
    if (p->bandpower && !(p->fused && p->fused_bp)) nm_describe_fir<NmEpiBandpower>(s, "bandpower", p->bandpower->bank, p->bandpower->epi_smem());
    if (p->sharpwave) nm_describe_fir<NmEpiSharpwave>(s, "sharpwave", p->sharpwave->bank, p->sharpwave->epi_smem());
    if (p->bursts) nm_describe_fir<NmEpiBursts>(s, "bursts", p->bursts->bank, p->bursts->epi_smem());
    if (buf && n > 0) {
        const size_t m = std::min<size_t>(s.size(), (size_t)n - 1);
        memcpy(buf, s.data(), m);
        buf[m] = 0;
    }
    return (int)s.size();
}
extern "C" int nm_synchronize(nm_pipeline* p) {
    NM_P_CHECK(p);
    cudaSetDevice(p->device);
    nm_upload_join(p);
    NM_CUDA_CHECK(cudaStreamSynchronize(p->copy_stream));
    NM_CUDA_CHECK(cudaStreamSynchronize(p->stream));
    return 0;
}
extern "C" int nm_result_device_ptr(nm_pipeline* p, void** ptr, long long* n_rows, int* n_cols) {
    NM_P_CHECK(p);
    if (ptr) *ptr = p->d_out.p;
    if (n_rows) *n_rows = p->out_rows;
    if (n_cols) *n_cols = p->F;
    return 0;
}
extern "C" int nm_stream_handle(nm_pipeline* p, void** cuda_stream) {
    NM_P_CHECK(p);
    if (cuda_stream) *cuda_stream = (void*)p->stream;
    return 0;
}

// ------------------------------------------------------------------------------- stand-alone FIR
// MNEFilter.filter_data (mode 0, filter/mne_filter.py:82-128) and NotchFilter.process (mode 1,
// filter/notch_filter.py:78-93) as one-shot calls for users of those classes outside a pipeline.
extern "C" int nm_fir_apply(int device, const double* taps, int n_filters, int n_taps, int mode, const double* data, int n_ch,
                            int n_samples, double* out) {
    NM_CHECK(taps && data && out && n_filters > 0 && n_taps > 0 && n_ch > 0 && n_samples >= 2, "bad arguments");
    NM_CHECK(mode == NM_FIR_SAME || mode == NM_FIR_REFLECT, "mode must be 0 (same) or 1 (reflect-limited)");
    int n_dev = 0;
    NM_CUDA_CHECK(cudaGetDeviceCount(&n_dev));
    NM_CHECK(n_dev > 0, "no CUDA device visible: libnmb200 has no CPU path");
    NM_CHECK(device >= 0 && device < n_dev, "device %d out of range", device);
    NM_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t s = nullptr;
    NM_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    int rc = -1;
    {
        FirBank bank;
        DevBuf d_in, d_out, d_off;
        nm_pipeline probe;  // only for the launch-geometry helpers
        cudaDeviceGetAttribute(&probe.n_sm, cudaDevAttrMultiProcessorCount, device);
        cudaDeviceGetAttribute(&probe.smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
        const long long zero = 0;
        const size_t sm_need = 0;
        (void)sm_need;
        do {
            if (bank.build(taps, n_filters, n_taps, n_samples, mode, s)) break;
            if (d_in.upload(data, (size_t)n_ch * n_samples, s)) break;
            if (d_off.upload(&zero, 1, s)) break;
            if (d_out.ensure((size_t)n_ch * n_filters * n_samples * sizeof(double))) break;
            NmRows rows;
            rows.base = d_in.as<double>();
            rows.ch_stride = n_samples;
            rows.off = d_off.as<long long>();
            rows.n_windows = 1;
            rows.n_ch = n_ch;
            rows.W = n_samples;
            NmEpiStore epi{d_out.as<double>(), (long long)n_samples, n_filters};
            if (nm_allow_fir_smem<NmEpiStore>(bank, 0, &probe)) break;
            nm_launch_fir(&probe, bank, rows, epi, s, 0);
            if (cudaGetLastError() != cudaSuccess) { nm_set_error("FIR kernel launch failed"); break; }
            if (cudaMemcpyAsync(out, d_out.p, (size_t)n_ch * n_filters * n_samples * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess) {
                nm_set_error("FIR result copy failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            rc = 0;
        } while (false);
    }
    cudaStreamDestroy(s);
    return rc;
}

// ------------------------------------------------------------------------------- sharded upload
// (implemented with the multi-GPU path; see nm_multi.cuh)
#include "nm_multi.cuh"
#include "nm_comm.cuh"
#include "nm_stream.cuh"
