// nm_stream.h -- state of the streaming entry (reference: stream/data_processor.py:238-311 called once per
// incoming window by stream/stream.py:280-330 / stream/mnelsl_stream.py).  Entry points: nm_stream.cuh.
//
//   * a RING of page-locked slots: the producer (an LSL reader, a file reader, the Python wrapper) writes the next window straight
//     into slot k's input block; nm_stream_submit(k) enqueues H2D + every kernel + D2H of the feature row and returns at once
//     ("async run_one"); nm_stream_wait(k) blocks on that slot's event only.  Slots let the transfer / kernels of window g + 1
//     overlap the host's consumption of window g.
//   * no host synchronisation inside a window: the per-window bookkeeping of the stateful families (burst-threshold quantile
//     positions, raw-normaliser history ranges) travels through a page-locked parameter block per slot instead of pageable
//     temporaries, the feature normaliser keeps a fixed-capacity right-aligned history (constant copy sizes).
//   * CUDA graph: the launch sequence of a window is captured once per slot and replayed with one cudaGraphLaunch; kernel
//     arguments that change from window to window (window counters) are patched into the executable graph (nm_platform.h).
#pragma once

#include <memory>
#include <vector>

#include "nm_host.h"

struct NmStreamSlot {
    void* in_host = nullptr;       // (C_all, Win) float32 / float64, page-locked
    double* out_host = nullptr;    // F doubles, page-locked
    unsigned char* par_host = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false, warm = false;
#ifndef NM_EMULATE
    NmGraphSession gs;
    cudaGraph_t graph = nullptr;
#endif
};

#define NM_STREAM_PAR_BYTES 4096

struct NmStream {
    int n_slots = 0;
    bool f32 = false, use_graph = false;
    std::vector<NmStreamSlot> slots;
    DevBuf d_par;                  // device copy of the current window's parameter block
    size_t par_used = 0;
    int cur = -1;                  // slot whose window is being enqueued (-1: not inside a streaming pass)
    long long windows = 0, graph_launches = 0;
};
