// nm_bursts.cuh -- burst features (features/bursts.py:149-298).
//
//  pass 1  nm_fir_kernel<NmEpiBursts>: band-pass bank ('same' FIR) -> W-point analytic-signal
//          envelope (scipy.signal.hilbert == W-point DFT, one-sided doubling, inverse DFT; the
//          Hilbert multiplier is real-linear, so the channel pair shares one complex transform)
//          -> envelope rows of the chunk + the newest samples appended to the history ring.
//  pass 2  nm_burst_thr_kernel: threshold = numpy 'linear' quantile of the last `ring` envelope
//          samples as of each window -- exact order statistics by radix select on the float64
//          bit patterns (envelopes are >= 0, so the patterns order like the values).
//  pass 3  nm_burst_feat_kernel: env >= thr, run-length statistics with warp ballots.
//
// The history is a TRUE ring of the last time_duration_s (the intended semantics; the reference's
// in-place partition scrambles its buffer once it is full -- SURVEY.md headline facts).  While the
// history is not full (windows 0..290 at defaults) both are identical.
#pragma once

#include "nm_fir.cuh"

struct NmEpiBursts {
    static constexpr bool kRegs = false;
    static constexpr bool kRegsOnly = false, kReflectOk = false, kSameOk = true, kConvxOnly = false;  // nm_convx_kernel instantiation traits
    NmFft<double> hfft;     // W-point transform
    int need_scratch;
    double* env;            // chunk envelopes (n_windows, n_ch, nB, Wp)
    long long Wp;
    int nB;
    double* ring;           // (n_ch, nB, cap)
    long long cap;
    long long win0;         // global index (since reset) of the chunk's first window
    int S;                  // samples appended by every window but the first
    static NM_HD size_t smem_bytes_for(int W, int need_scratch) { return (size_t)W * sizeof(cx<double>) * (need_scratch ? 2 : 1); }

    NM_DEV void run(const cx<double>* buf, int o0, int W, int n_ch, int w, int c0, bool has2, int f,
                    unsigned char* scratch_raw, int tid, int nt) const {
        cx<double>* hb = reinterpret_cast<cx<double>*>(scratch_raw);
        cx<double>* sc = need_scratch ? hb + W : nullptr;
        const cx<double>* x = buf + o0;
        for (int t = tid; t < W; t += nt) hb[t] = x[t];
        __syncthreads();
        nm_fft_forward<double>(hb, sc, hfft, tid, nt);
        const double inv = 1.0 / W;
        for (int k = tid; k < W; k += nt) {
            const int slot = nm_ldg(hfft.pos + k);
            const cx<double> v = hb[slot];
            cx<double> r = {0.0, 0.0};
            if (k != 0 && 2 * k != W) {
                if (2 * k < W) r = {v.im * inv, -v.re * inv};   // * (-i)
                else r = {-v.im * inv, v.re * inv};             // * (+i)
            }
            hb[slot] = r;
        }
        __syncthreads();
        nm_fft_inverse<double>(hb, sc, hfft, tid, nt);
        const long long gw = win0 + w;
        const int take = (gw == 0) ? W : S;
        const long long e_prev = (gw == 0) ? 0 : (long long)W + (gw - 1) * S;
        for (int k = 0; k < (has2 ? 2 : 1); ++k) {
            const int c = c0 + k;
            double* erow = env + (((size_t)w * n_ch + c) * nB + f) * Wp;
            double* rrow = ring + ((size_t)c * nB + f) * cap;
            for (int t = tid; t < W; t += nt) {
                const double re = k ? x[t].im : x[t].re;
                const double im = k ? hb[t].im : hb[t].re;
                const double e = sqrt(re * re + im * im);
                erow[t] = e;
                const int j = t - (W - take);
                if (j >= 0) rrow[(e_prev + j) % cap] = e;
            }
        }
    }
};

// ------------------------------------------------------------------ pass 2: thresholds
struct NmBurstThrArgs {
    const double* ring;
    long long cap;
    int n_ch, nB, n_windows;
    const long long* e_end;  // [n_windows] logical end (exclusive) of the history as of window w
    const int* n_hist;       // [n_windows] number of samples in the history
    const int* k_lo;         // [n_windows] order statistic below the virtual index
    const int* k_hi;
    const double* gamma;     // [n_windows]
    double* thr;             // (n_windows, n_ch, nB)
};

#define NM_SEL_BINS 4096
#define NM_SEL_CAND 1024

NM_GLOBAL void nm_burst_thr_kernel(NmBurstThrArgs a) {
    NM_SHARED_BYTES(smem);
    int* hist = reinterpret_cast<int*>(smem);                                 // NM_SEL_BINS
    unsigned long long* cand = reinterpret_cast<unsigned long long*>(hist + NM_SEL_BINS);  // NM_SEL_CAND
    int* ctl = reinterpret_cast<int*>(cand + NM_SEL_CAND);                    // [0] digit [1] below [2] count [3] ncand
    unsigned long long* res = reinterpret_cast<unsigned long long*>(ctl + 8); // [0] a  [1] min greater
    int* cnts = reinterpret_cast<int*>(res + 2);                              // [0] less [1] equal
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;

    const long long n_items = (long long)a.n_windows * a.n_ch * a.nB;
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int w = (int)(item / ((long long)a.n_ch * a.nB));
        const int cb = (int)(item - (long long)w * a.n_ch * a.nB);
        const double* rrow = a.ring + (size_t)cb * a.cap;
        const int n = a.n_hist[w];
        const long long first = a.e_end[w] - n;
        const long long first_mod = first % a.cap;
        int rank = a.k_lo[w];
        unsigned long long prefix = 0ull, mask = 0ull;
        int shift = 52, width = 12;
        bool resolved = false;
        while (true) {
            for (int i = tid; i < NM_SEL_BINS; i += nt) hist[i] = 0;
            __syncthreads();
            const unsigned long long dm = (1ull << width) - 1ull;
            for (int i = tid; i < n; i += nt) {
                long long p = first_mod + i;
                if (p >= a.cap) p -= a.cap;
                const unsigned long long key = (unsigned long long)__double_as_longlong(rrow[p]);
                if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & dm)], 1);
            }
            __syncthreads();
            if (tid < 32) {
                const int per = NM_SEL_BINS / 32;
                int s = 0;
                for (int i = 0; i < per; ++i) s += hist[lane * per + i];
                int incl = s;
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                const int excl = incl - s;
                if (rank >= excl && rank < incl) {
                    int run = excl;
                    for (int i = 0; i < per; ++i) {
                        const int h = hist[lane * per + i];
                        if (rank < run + h) {
                            ctl[0] = lane * per + i;
                            ctl[1] = run;
                            ctl[2] = h;
                            break;
                        }
                        run += h;
                    }
                }
            }
            __syncthreads();
            const int digit = ctl[0], below = ctl[1], cnt = ctl[2];
            rank -= below;
            prefix |= ((unsigned long long)digit) << shift;
            mask |= dm << shift;
            __syncthreads();
            if (shift == 0) { resolved = true; break; }
            if (cnt <= NM_SEL_CAND) break;
            if (shift >= 16) { shift -= 12; width = 12; }
            else { width = shift; shift = 0; }   // final 4 bits
        }
        unsigned long long a_key = prefix;
        if (!resolved) {
            if (tid == 0) ctl[3] = 0;
            __syncthreads();
            for (int i = tid; i < n; i += nt) {
                long long p = first_mod + i;
                if (p >= a.cap) p -= a.cap;
                const unsigned long long key = (unsigned long long)__double_as_longlong(rrow[p]);
                if ((key & mask) == prefix) {
                    const int slot = atomicAdd(&ctl[3], 1);
                    if (slot < NM_SEL_CAND) cand[slot] = key;
                }
            }
            __syncthreads();
            const int m = ctl[3] < NM_SEL_CAND ? ctl[3] : NM_SEL_CAND;
            for (int i = tid; i < m; i += nt) {
                const unsigned long long x = cand[i];
                int r = 0;
                for (int j = 0; j < m; ++j) {
                    const unsigned long long y = cand[j];
                    r += (y < x || (y == x && j < i)) ? 1 : 0;
                }
                if (r == rank) res[0] = x;
            }
            __syncthreads();
            a_key = res[0];
        }
        // second order statistic: a again if duplicated far enough, else the smallest larger value
        double thr;
        const double av = __longlong_as_double((long long)a_key);
        if (a.k_hi[w] == a.k_lo[w]) {
            thr = av;
        } else {
            if (tid == 0) { cnts[0] = 0; cnts[1] = 0; res[1] = ~0ull; }
            __syncthreads();
            int less = 0, eq = 0;
            unsigned long long mg = ~0ull;
            for (int i = tid; i < n; i += nt) {
                long long p = first_mod + i;
                if (p >= a.cap) p -= a.cap;
                const unsigned long long key = (unsigned long long)__double_as_longlong(rrow[p]);
                less += key < a_key;
                eq += key == a_key;
                if (key > a_key && key < mg) mg = key;
            }
            less = nm_warp_sum_i(less);
            eq = nm_warp_sum_i(eq);
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long u = __shfl_xor_sync(0xffffffffu, mg, o);
                mg = u < mg ? u : mg;
            }
            if (lane == 0) {
                atomicAdd(&cnts[0], less);
                atomicAdd(&cnts[1], eq);
                // 64-bit min via two-step: serialised by warp count (<= 32 warps), done with a spin-free CAS-less loop below
            }
            __syncthreads();
            // reduce the per-warp minima through shared memory (cand[] is free now)
            if (lane == 0) cand[tid >> 5] = mg;
            __syncthreads();
            if (tid == 0) {
                unsigned long long best = ~0ull;
                for (int q = 0; q < ((nt + 31) >> 5); ++q) best = cand[q] < best ? cand[q] : best;
                res[1] = best;
            }
            __syncthreads();
            const double bv = (a.k_hi[w] < cnts[0] + cnts[1]) ? av : __longlong_as_double((long long)res[1]);
            const double g = a.gamma[w];
            const double diff = bv - av;
            double lerp = av + diff * g;
            if (g >= 0.5) lerp = bv - diff * (1.0 - g);
            thr = (diff == 0.0) ? av : lerp;
        }
        if (tid == 0) a.thr[item] = thr;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ pass 3: run-length features
struct NmBurstFeatArgs {
    const double* env;   // (n_windows, n_ch, nB, Wp)
    long long Wp;
    const double* thr;   // (n_windows, n_ch, nB)
    int n_windows, n_ch, nB, W;
    double sfreq, seg_s;
    NmOut out;           // per_ch = nB * 6
};

NM_DEV double nm_warp_incl_scan(double v, int lane) {
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

NM_GLOBAL void nm_burst_feat_kernel(NmBurstFeatArgs a) {
    const int lane = threadIdx.x & 31;
    const int wpc = blockDim.x >> 5;
    const long long n_rows = (long long)a.n_windows * a.n_ch * a.nB;
    const int W = a.W;
    for (long long row = (long long)blockIdx.x * wpc + (threadIdx.x >> 5); row < n_rows; row += (long long)gridDim.x * wpc) {
        const double* e = a.env + (size_t)row * a.Wp;
        const double thr = a.thr[row];
        int total = 0, trans = 0, nv = 0, maxlen = 0;
        double means = 0.0, amax = 0.0;
        int open_len = 0;
        double open_sum = 0.0;
        unsigned prev_last = 0u;
        const int nsteps = (W + 31) >> 5;
        for (int j = 0; j < nsteps; ++j) {
            const int t = (j << 5) + lane;
            const bool valid = t < W;
            const double ev = valid ? e[t] : 0.0;
            const bool b = valid && (ev >= thr);
            const unsigned mask = __ballot_sync(0xffffffffu, b);
            const unsigned vmask = (j == nsteps - 1 && (W & 31)) ? ((1u << (W & 31)) - 1u) : 0xffffffffu;
            total += __popc(mask);
            trans += __popc((mask ^ ((mask << 1) | prev_last)) & vmask);
            if (b && ev > amax) amax = ev;
            const double P = nm_warp_incl_scan(b ? ev : 0.0, lane);
            // a run carried over from the previous step that stops exactly at the step boundary
            if (open_len > 0 && !(mask & 1u)) {
                if (lane == 0) {
                    nv += 1;
                    maxlen = open_len > maxlen ? open_len : maxlen;
                    means += open_sum / open_len;
                }
                open_len = 0;
                open_sum = 0.0;
            }
            // run bookkeeping for lanes inside a run
            const unsigned below = ~mask & ((lane == 0) ? 0u : (0xffffffffu >> (32 - lane)));
            const int seg_start = below ? (32 - __clz((int)below)) : 0;
            const bool from_carry = (below == 0u);
            const double Pbefore = __shfl_sync(0xffffffffu, P, seg_start > 0 ? seg_start - 1 : 0);
            const double runsum = P - (seg_start > 0 ? Pbefore : 0.0);
            const bool is_end = b && lane < 31 && !((mask >> (lane + 1)) & 1u);
            if (is_end) {
                const int len = (lane - seg_start + 1) + (from_carry ? open_len : 0);
                const double sum = runsum + (from_carry ? open_sum : 0.0);
                if (t < W - 1) {
                    nv += 1;
                    maxlen = len > maxlen ? len : maxlen;
                    means += sum / len;
                }
            }
            // carry for the next step (warp-uniform: taken from lane 31)
            const int l31_start = __shfl_sync(0xffffffffu, seg_start, 31);
            const int l31_carry = __shfl_sync(0xffffffffu, (int)from_carry, 31);
            const double l31_sum = __shfl_sync(0xffffffffu, runsum, 31);
            if (mask >> 31) {
                if (l31_carry) { open_len += 32; open_sum += l31_sum; }
                else { open_len = 32 - l31_start; open_sum = l31_sum; }
            } else {
                open_len = 0;
                open_sum = 0.0;
            }
            prev_last = mask >> 31;
        }
        nv = nm_warp_sum_i(nv);
        means = nm_warp_sum(means);
        amax = nm_warp_max(amax);
        for (int o = 16; o > 0; o >>= 1) {
            const int u = __shfl_xor_sync(0xffffffffu, maxlen, o);
            maxlen = u > maxlen ? u : maxlen;
        }
        if (lane == 0) {
            const int w = (int)(row / ((long long)a.n_ch * a.nB));
            const int cb = (int)(row - (long long)w * a.n_ch * a.nB);
            const int c = cb / a.nB, bnd = cb - c * a.nB;
            const int num = trans / 2;
            const double dmean = num ? ((double)total / (double)num) / a.sfreq : 0.0;
            nm_store(a.out, w, c, bnd * 6 + 0, dmean);
            nm_store(a.out, w, c, bnd * 6 + 1, (double)maxlen / a.sfreq);
            nm_store(a.out, w, c, bnd * 6 + 2, nv ? means / nv : 0.0);
            nm_store(a.out, w, c, bnd * 6 + 3, amax);
            nm_store(a.out, w, c, bnd * 6 + 4, dmean / a.seg_s);
            nm_store(a.out, w, c, bnd * 6 + 5, (double)((thr <= e[W - 1]) ? 1 : 0));
        }
    }
}

// ------------------------------------------------------------------ host side
struct BurstsFam {
    FirBank bank;
    FftPlanHost hfft;
    DevBuf d_env, d_ring, d_thr, d_colmap, d_e_end, d_n, d_lo, d_hi, d_gamma;
    int nB = 0, C = 0, W = 0, S = 0, ring_n = 0, chunk = 0;
    long long cap = 0, Wp = 0, batch = 0;
    double q = 0.75, sfreq = 1000, seg_s = 1;

    int build(const double* taps, int nB_, int L, int C_, int W_, int S_, int ring_n_, double q_, double sfreq_, double seg_s_,
              const int* colmap, cudaStream_t s) {
        nB = nB_; C = C_; W = W_; S = S_; ring_n = ring_n_; q = q_; sfreq = sfreq_; seg_s = seg_s_;
        if (bank.build(taps, nB, L, W, NM_FIR_SAME, s)) return -1;
        if (hfft.build(W, s)) return -1;
        return d_colmap.upload(colmap, (size_t)C * nB * 6, s);
    }
    int alloc_chunk(int chunk_, long long Wp_) {
        chunk = chunk_;
        Wp = Wp_;
        cap = (long long)ring_n + (long long)chunk * (S > 0 ? S : 1) + W;
        if (d_env.ensure((size_t)chunk * C * nB * Wp * sizeof(double))) return -1;
        if (d_ring.ensure((size_t)C * nB * cap * sizeof(double))) return -1;
        if (d_thr.ensure((size_t)chunk * C * nB * sizeof(double))) return -1;
        return 0;
    }
    void reset() { batch = 0; }
    size_t epi_smem() const { return NmEpiBursts::smem_bytes_for(W, hfft.generic); }
    static size_t thr_smem() { return NM_SEL_BINS * sizeof(int) + NM_SEL_CAND * 8 + 8 * sizeof(int) + 2 * 8 + 2 * sizeof(int) + 64; }
    int allow_smem(const nm_pipeline* p);
    int run(nm_pipeline* p, const NmRows& rows, int w0);
};
