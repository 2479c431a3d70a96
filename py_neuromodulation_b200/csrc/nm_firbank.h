// nm_firbank.h -- host-side description of one FIR bank (taps -> transform size, spectra, kernel args).
#pragma once

#include <algorithm>

#include "nm_fir.cuh"
#include "nm_host.h"

struct nm_pipeline;

struct FirBank {
    int nF = 0, L = 0, Lh = 0, P = 0, mode = NM_FIR_SAME, E = 0;
    FftPlanHost fft;
    DevBuf d_hperm;
    int build(const double* taps, int nF_, int L_, int W, int mode_, cudaStream_t s) {
        nF = nF_; L = L_; Lh = (L - 1) / 2; mode = mode_;
        if (mode == NM_FIR_REFLECT) {
            // mne _overlap_add_filter: n_edge = max(min(len(h), len(x)) - 1, 0) reflected samples per side;
            // only the (L-1)/2 nearest ones can reach the W centre outputs
            const int n_edge = std::max(std::min(L, W) - 1, 0);
            E = std::min(n_edge, Lh);
            P = nm_next_smooth(W + E + Lh);
        } else {
            E = 0;
            P = nm_next_smooth(W + Lh);
        }
        if (fft.build(P, s)) return -1;
        NM_CHECK(!fft.generic, "internal: convolution length %d is not 5-smooth", P);
        std::vector<double> hperm;
        if (nm_build_hperm(taps, nF, L, fft, hperm)) return -1;
        return d_hperm.upload(hperm, s);
    }
    NmFirArgs args(const NmRows& in) const {
        NmFirArgs a;
        a.in = in;
        a.fft = fft.dev();
        a.hperm = d_hperm.as<double>();
        a.nF = nF;
        a.mode = mode;
        a.E = E;
        a.n_items = in.n_windows * ((in.n_ch + 1) / 2);
        return a;
    }
    size_t smem(size_t epi) const { return (size_t)P * sizeof(cx<double>) * (nF > 1 ? 2 : 1) + epi; }
};
