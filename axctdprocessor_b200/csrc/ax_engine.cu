// ax_engine.cu -- kernels, host runtime and C ABI of the AXCTD engine (sm_100a).
//
// Built by __graft_entry__.build() with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
// The same translation unit compiles with g++ -x c++ -DAXCTD_EMU into the
// test-only host emulation (tests/emu), which the product never loads.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <string>
#include <vector>

#include "ax_proto.h"
#include "ax_synth.h"

#ifndef AXCTD_EMU
#include <cuda_runtime.h>
#include "ax_kernels.cuh"
#endif

// ============================================================ launch plumbing
#ifdef AXCTD_EMU
typedef int axStream;
#define AX_GLOBAL static
#define AX_FOR_ITEM(n) for (int64_t item = 0; item < (n); ++item)
#define AX_FOR_ITEM1(n) for (int64_t item = 0; item < (n); ++item)
#define AX_LAUNCH(eng, name, n, ...)                         \
    do { if ((n) > 0) { name((int64_t)(n), __VA_ARGS__); (eng)->launches++; } } while (0)
#define AX_LAUNCH1 AX_LAUNCH
#else
typedef cudaStream_t axStream;
#define AX_GLOBAL __global__
#define AX_FOR_ITEM(n) const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (item < (n))
#define AX_LAUNCH(eng, name, n, ...)                                                                   \
    do { if ((n) > 0) { const int64_t _n = (n); const int _b = 128;                                   \
        name<<<(unsigned)((_n + _b - 1) / _b), _b, 0, (eng)->stream>>>(_n, __VA_ARGS__); (eng)->launches++; } } while (0)
// sequential per-drop work: one item per CTA (its own SM, no divergence between the items of a warp)
#define AX_FOR_ITEM1(n) const int64_t item = blockIdx.x; if (threadIdx.x == 0 && item < (n))
#define AX_LAUNCH1(eng, name, n, ...)                                                                  \
    do { if ((n) > 0) { name<<<(unsigned)(n), 32, 0, (eng)->stream>>>((int64_t)(n), __VA_ARGS__); (eng)->launches++; } } while (0)
#endif

AX_GLOBAL void k_init(int64_t n, AxWave w) {
    AX_FOR_ITEM(n) {
        AxState& st = w.st[item];
        memset(&st, 0, sizeof(AxState));
        ax_state_defaults(st, w.cfg[w.drop[item].cfg]);
    }
}
// ---- streaming decode of a growing recording (axctd_batch_stream_*)
// normalisation fixed by the caller instead of the whole-file statistics of AXCTDprocessor.py:55-57 (ax_stats_fin)
AX_GLOBAL void k_stream_begin(int64_t n, AxWave w, const double* norm) {
    AX_FOR_ITEM(n) {
        AxState& st = w.st[item];
        st.dc = norm[2 * item]; st.ampl_d = norm[2 * item + 1];
        st.ampl = (int32_t)st.ampl_d; st.inv_ampl = ax_div(1.0, st.ampl_d);
    }
}
// a later run: the state of the iterations already decoded stays; only what a run re-derives is re-armed
AX_GLOBAL void k_stream_resume(int64_t n, AxWave w) {
    AX_FOR_ITEM(n) {
        AxState& st = w.st[item];
        st.chain_end = 0; st.chain_dirty = 0;
        st.n_unc_relevant = 0; st.n_unc_resolved = 0; st.n_uncertain = 0;      // (ax_unc_resolve_item goes through the whole list again)
    }
}
// end of a run: what it leaves final for the next one
AX_GLOBAL void k_stream_commit(int64_t n, AxWave w) {
    AX_FOR_ITEM(n) {
        AxState& st = w.st[item];
        const AxDrop& dr = w.drop[item];
        const AxCfg& c = w.cfg[dr.cfg];
        if (st.status == 0) st.k_done = st.n_chunks;
        int64_t sd = dr.n > c.npcm ? (dr.n - c.npcm) / w.seg_len : 0;        // segments whose last window lies inside the data
        if (sd > dr.nseg) sd = dr.nseg;
        st.seg_done = (int32_t)sd;
        st.tb_done = dr.ntb;
    }
}
AX_GLOBAL void k_inject(int64_t n, AxWave w) {     // test hook: pretend the first prediction was off by 3 samples
    AX_FOR_ITEM(n) {
        AxState& s = w.st[item];
        if (s.sm_status >= 1 && s.n_chunks > s.k0 + 2) {
            AxChunk* c = w.chunk + w.drop[item].chunk_base + s.k0;
            c[0].spec_last += 3; c[1].s += 3; c[1].e += 3;
        }
    }
}
AX_GLOBAL void k_stats(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_stats_item(w, item); }
AX_GLOBAL void k_stats_fin(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_stats_fin(w, item); }
AX_GLOBAL void k_filter(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_filter_item(w, item); }
AX_GLOBAL void k_scan_block(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_scan_block_item(w, item); }
AX_GLOBAL void k_scan(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_scan_item(w, item); }
AX_GLOBAL void k_compact(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_compact_item(w, item); }
AX_GLOBAL void k_nx(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_nx_item(w, item); }
AX_GLOBAL void k_tiles(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_tiles_item(w, item); }
AX_GLOBAL void k_plan0(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_plan0_item(w, item); }
AX_GLOBAL void k_tone_direct(int64_t n, AxWave w, int phase_b) { AX_FOR_ITEM(n) ax_tone_direct_item(w, item, phase_b); }
AX_GLOBAL void k_toneblock(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_toneblock_item(w, item); }
AX_GLOBAL void k_tonewin(int64_t n, AxWave w, int phase_b) { AX_FOR_ITEM(n) ax_tonewin_item(w, item, phase_b); }
AX_GLOBAL void k_pwfill(int64_t n, AxWave w, int phase_b) { AX_FOR_ITEM(n) ax_pwfill_item(w, item, phase_b); }
AX_GLOBAL void k_levels(int64_t n, AxWave w, int phase_b) { AX_FOR_ITEM(n) ax_levels_item(w, item, phase_b); }
AX_GLOBAL void k_sm(int64_t n, AxWave w, int phase_b) { AX_FOR_ITEM1(n) ax_sm_item(w, item, phase_b); }
AX_GLOBAL void k_canon(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_canon_item(w, item); }
AX_GLOBAL void k_chain(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_chain_item(w, item); }
AX_GLOBAL void k_headfilt(int64_t n, AxWave w, int only_rest) { AX_FOR_ITEM(n) ax_headfilt_item(w, item, only_rest); }
AX_GLOBAL void k_headwalk(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_headwalk_item(w, item); }
AX_GLOBAL void k_verify(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_verify_item(w, item); }
AX_GLOBAL void k_unc_resolve(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_unc_resolve_item(w, item); }
AX_GLOBAL void k_unc_fin(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_unc_fin_item(w, item); }
AX_GLOBAL void k_plan_tones(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_plan_tones_item(w, item); }
AX_GLOBAL void k_offsets(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_offsets_item(w, item); }
AX_GLOBAL void k_emit(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_emit_item(w, item); }
AX_GLOBAL void k_scale(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_scale_item(w, item); }
AX_GLOBAL void k_bits(int64_t n, AxWave w, int phase) { AX_FOR_ITEM(n) ax_bits_item(w, item, phase); }
AX_GLOBAL void k_headers(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_header_item(w, item); }
AX_GLOBAL void k_merge(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_merge_item(w, item); }
AX_GLOBAL void k_pack(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_pack_item(w, item); }
AX_GLOBAL void k_valid(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_valid_item(w, item); }
AX_GLOBAL void k_frames_spec(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_frames_spec_item(w, item); }
AX_GLOBAL void k_frames_chain(int64_t n, AxWave w) { AX_FOR_ITEM1(n) ax_frames_chain_item(w, item); }
AX_GLOBAL void k_frames_write(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_frames_write_item(w, item); }
AX_GLOBAL void k_calib(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_calib_item(w, item); }
AX_GLOBAL void k_synth(int64_t n, AxSynth g) { AX_FOR_ITEM(n) ax_synth_item(g, item); }
AX_GLOBAL void k_qc(int64_t n, AxWave w, double* scratch) { AX_FOR_ITEM(n) ax_qc_item(w, item, scratch); }
AX_GLOBAL void k_decim(int64_t n, AxWave w, int pass) { AX_FOR_ITEM(n) ax_decim_item(w, item, pass); }
AX_GLOBAL void k_decim_fin(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_decim_fin(w, item); }
AX_GLOBAL void k_calib_eval(int64_t n, const double* in, double* out, int has_coeff) {
    AX_FOR_ITEM(n) {       // in: [3][n] cond, temp, pres then 4 coefficients; out: [2][n] sp, poly
        out[item] = ax_sp_from_c(in[item], in[n + item], in[2 * n + item]);
        out[n + item] = has_coeff ? ax_dataconvert(in[item], in + 3 * n) : 0.0;
    }
}
// first channel of interleaved frames (AXCTDprocessor.py:50): eight output samples per item, 16-byte loads and stores
// when the frames allow it
AX_GLOBAL void k_deinterleave(int64_t n, const int16_t* frames, int16_t* out, int64_t n_frames, int channels) {
    AX_FOR_ITEM(n) {
        const int64_t f0 = item * 8;
        int16_t v[8];
#if defined(__CUDA_ARCH__)
        if (channels == 2 && f0 + 8 <= n_frames) {
            const int4 a = reinterpret_cast<const int4*>(frames)[2 * item], c = reinterpret_cast<const int4*>(frames)[2 * item + 1];
            int4 o;
            o.x = __byte_perm(a.x, a.y, 0x5410); o.y = __byte_perm(a.z, a.w, 0x5410);
            o.z = __byte_perm(c.x, c.y, 0x5410); o.w = __byte_perm(c.z, c.w, 0x5410);
            reinterpret_cast<int4*>(out)[item] = o;
        } else
#endif
        {
            for (int q = 0; q < 8; ++q) v[q] = f0 + q < n_frames ? frames[(f0 + q) * channels] : (int16_t)0;
            for (int q = 0; q < 8 && f0 + q < n_frames; ++q) out[f0 + q] = v[q];
        }
    }
}
AX_GLOBAL void k_rows(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_row_item(w, item); }
AX_GLOBAL void k_chunkout(int64_t n, AxWave w) { AX_FOR_ITEM(n) ax_chunkout_item(w, item); }

// ============================================================ memory helpers
struct axctd_engine {
    int device = 0;
    axStream stream = 0;
    bool own_stream = true;
#ifndef AXCTD_EMU
    cudaStream_t hp_stream = nullptr;     // highest priority: the demodulation pass gets the SMs first when sub-batches share the GPU
#endif
    int opt_heavy_prio = 1;
    std::string err;
    std::vector<AxCfg> cfgs;              // host copies (device pointers inside)
    std::vector<AxToneTab> tone_tabs;     // per config: phasors of one tone block (kernel parameter)
    std::vector<void*> cfg_allocs;
    AxCfg* d_cfg = nullptr;
    int cfg_cap = 64;
    int64_t launches = 0;
    // options
    int64_t opt_segment_len = 0;          // 0 = auto
    int opt_exact_head = 0;               // 0 = auto from pole radius
    double opt_guard = 1e-12;
    int opt_force_exact = 0;
    int opt_tone_direct = 0;
    int opt_max_fixups = 64;
    int opt_filter_variant = 0;
    int opt_zc_div = 12;
    int opt_inject_misspec = 0;           // test hook: corrupt the first prediction
    double opt_bit_tol = 2e-5;            // fp32 bit windows: relative distance to a decision boundary that triggers
    double opt_hist_tol = 2e-5;           //   the double-precision re-evaluation (bit decision / calibration histogram)
    int opt_bitfix_all = 0;               // test hook: re-evaluate every window
    int opt_demod_probe = 0;              // timing probe (AxWave::probe)
    int opt_ws = 0;                       // warp-specialised fused kernel (k_demod_ws)
    int opt_bulk = 0;                     // continuous pass stages its rows with cp.async.bulk (TMA unit) instead of LDGSTS
    int opt_fir_first = 1;                // numerators-first cascade in the continuous low-pass pass (k_demod_fused FAST)
    int opt_tone_mma = 1;                 // tone block sums on the FP64 tensor cores (k_stats_tones_mma)
    int opt_seg_target = 16384;           // segment length the number of waves of the demodulation pass is sized for (8192: 15.3 ms per step, 16384: 14.7)
    int opt_tone_complement = 1;          // ragged window ends above half a tone block as the block minus its complement (k_tone_windows_mma)
    int opt_tone_int8 = 1;                // ... as exact integer products on the int8 tensor cores instead (k_stats_tones_imma); 0: FP64 tensor cores
    int opt_pair_launch = 0;              // two rate classes (window lengths 39 / 43) demodulated by one launch (k_demod_fused_pair): measured, no gain
    int opt_heavy_chain = 1;              // engines of one process take turns with the demodulation pass (see ax_heavy_*)
    int opt_nosync = 1;                   // enqueue the whole decode without host round trips (see axctd_batch_run_async)
    int opt_scan_only = 0;                // tone levels only (segmentation of long recordings): skip the demodulation pass
    int opt_fuse_bits = 1;                // bits decided while the edges are emitted (k_emit_chunk mode 2)
};

#ifdef AXCTD_EMU
static int ax_alloc(axctd_engine*, void** p, size_t bytes) { *p = calloc(bytes ? bytes : 1, 1); return *p ? 0 : 1; }
static void ax_free(axctd_engine*, void* p) { free(p); }
static int ax_h2d(axctd_engine*, void* d, const void* h, size_t b) { memcpy(d, h, b); return 0; }
static int ax_d2h(axctd_engine*, void* h, const void* d, size_t b) { memcpy(h, d, b); return 0; }
static int ax_zero(axctd_engine*, void* d, size_t b) { memset(d, 0, b); return 0; }
static int ax_sync(axctd_engine*) { return 0; }
static int ax_launch_check(axctd_engine*) { return 0; }
#define AX_DEV(e) ((void)0)
#else
static int ax_fail(axctd_engine* e, cudaError_t r, const char* what) {
    if (r == cudaSuccess) return 0;
    e->err = std::string(what) + ": " + cudaGetErrorString(r);
    return 1;
}
// ---- block cache.  cudaMalloc / cudaFree / cudaHostAlloc of the gigabyte-sized arrays of a batch cost more than the
// decode of a one-hour recording (measured: 0.09 s to create and 0.58 s to destroy the two batches of BASELINE config 3
// against 0.045 s of device work, profiles/r2_run5_config3_phases_before.json), and cudaFree synchronises the whole
// device.  Blocks of destroyed batches are therefore kept, per device (one list for pinned host memory), and handed
// to the next batch that asks for about that size; the cache is emptied when the last engine of the process on that
// device is destroyed, when an allocation fails, or on request (engine option "pool_trim").  A block only enters the
// cache after its batch's stream has been synchronised (axctd_batch_destroy).
#include <map>
#include <mutex>
#include <unordered_map>
struct AxBlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> idle;             // cached blocks by size
    std::unordered_map<void*, size_t> live;        // size of every block handed out
    size_t idle_bytes = 0;
    int engines = 0;                               // live engines using this cache
    void* take(size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        auto it = idle.lower_bound(bytes);
        if (it == idle.end() || it->first > bytes + bytes / 4 + 65536) return nullptr;
        void* p = it->second;
        live[p] = it->first; idle_bytes -= it->first;
        idle.erase(it);
        return p;
    }
    void adopt(void* p, size_t bytes) { std::lock_guard<std::mutex> g(mu); live[p] = bytes; }
    bool give(void* p) {                           // false: not one of ours
        std::lock_guard<std::mutex> g(mu);
        auto it = live.find(p);
        if (it == live.end()) return false;
        idle.emplace(it->second, p); idle_bytes += it->second;
        live.erase(it);
        return true;
    }
    template <typename F> void trim(F release) {
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : idle) release(kv.second);
        idle.clear(); idle_bytes = 0;
    }
};
static AxBlockCache g_dev_cache[64];
static AxBlockCache g_host_cache;
static int g_pool_on = 1, g_pool_poison = 0;
static AxBlockCache* ax_dev_cache(const axctd_engine* e) { return (g_pool_on && e->device >= 0 && e->device < 64) ? &g_dev_cache[e->device] : nullptr; }
static void ax_pool_trim(int device) {
    if (device >= 0 && device < 64) g_dev_cache[device].trim([](void* p) { cudaFree(p); });
    g_host_cache.trim([](void* p) { cudaFreeHost(p); });
}
static int ax_alloc(axctd_engine* e, void** p, size_t bytes) {
    bytes = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
    AxBlockCache* c = ax_dev_cache(e);
    if (c && (*p = c->take(bytes)) != nullptr) {
        if (g_pool_poison) cudaMemsetAsync(*p, 0xA5, bytes, e->stream);      // test hook: nothing may rely on fresh (zeroed) pages
        return 0;
    }
    cudaError_t r = cudaMalloc(p, bytes);
    if (r == cudaErrorMemoryAllocation) { cudaGetLastError(); ax_pool_trim(e->device); r = cudaMalloc(p, bytes); }
    if (r == cudaSuccess && c) c->adopt(*p, bytes);
    return ax_fail(e, r, "cudaMalloc");
}
static void ax_free(axctd_engine* e, void* p) {
    if (!p) return;
    AxBlockCache* c = ax_dev_cache(e);
    if (c && c->give(p)) return;
    if (e->device >= 0 && e->device < 64) { std::lock_guard<std::mutex> g(g_dev_cache[e->device].mu); g_dev_cache[e->device].live.erase(p); }
    cudaFree(p);
}
static int ax_h2d(axctd_engine* e, void* d, const void* h, size_t b) { return ax_fail(e, cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, e->stream), "H2D"); }
static int ax_d2h(axctd_engine* e, void* h, const void* d, size_t b) { return ax_fail(e, cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, e->stream), "D2H"); }
static int ax_zero(axctd_engine* e, void* d, size_t b) { return ax_fail(e, cudaMemsetAsync(d, 0, b, e->stream), "memset"); }
// A bad launch configuration (missing shared-memory opt-in, oversized grid ...) is reported by cudaGetLastError only,
// never by a later stream synchronisation: every sync point looks at it first, so that a failed launch fails the batch
// instead of leaving initial values in the results.
static int ax_launch_check(axctd_engine* e) { return ax_fail(e, cudaGetLastError(), "kernel launch"); }
static int ax_sync(axctd_engine* e) {
    if (ax_launch_check(e)) return 1;
    return ax_fail(e, cudaStreamSynchronize(e->stream), "sync");
}
#define AX_DEV(e) cudaSetDevice((e)->device)          /* every ABI entry point that touches CUDA runs on its engine's device */
#endif

#ifdef AXCTD_EMU
static void* ax_host_alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
static void ax_host_free(void* p) { free(p); }
#else
static void* ax_host_alloc(size_t bytes) {          // pinned, through the block cache
    bytes = ((bytes ? bytes : 1) + 4095) & ~(size_t)4095;
    void* p = g_pool_on ? g_host_cache.take(bytes) : nullptr;
    if (p) return p;
    cudaError_t r = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (r != cudaSuccess) { cudaGetLastError(); g_host_cache.trim([](void* q) { cudaFreeHost(q); }); r = cudaHostAlloc(&p, bytes, cudaHostAllocDefault); }
    if (r != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (g_pool_on) g_host_cache.adopt(p, bytes);
    return p;
}
static void ax_host_free(void* p) {
    if (!p) return;
    if (g_pool_on && g_host_cache.give(p)) return;
    { std::lock_guard<std::mutex> g(g_host_cache.mu); g_host_cache.live.erase(p); }
    cudaFreeHost(p);
}
#endif

struct axctd_batch {
    axctd_engine* eng = nullptr;
    int n = 0;
    std::vector<AxDrop> drops;
    std::vector<void*> allocs;
    AxWave w;
    int16_t* d_pcm = nullptr;
    void* d_stage = nullptr; size_t stage_bytes = 0;      // interleaved frames of a multi-channel upload
    double* d_qc = nullptr;
    int64_t tb_total = 0, dseg_total = 0;
    int64_t pcm_total = 0, chunk_total = 0, edge_total = 0, frame_total = 0, zc_total = 0, tile_total = 0;
    // host mirrors of the results (pinned: the result download is part of every step)
    std::vector<AxState> st;
    AxState* h_st = nullptr;
    axctd_row* h_row = nullptr;           // [frame_total], a drop's rows at frame_base
    axctd_chunk* h_chunk = nullptr;       // [chunk_total]
    std::vector<axctd_drop_summary> summary;
    bool force_sync = false;              // repeat of a run whose fixed schedule did not suffice
    bool force_nofuse = false;            // repeat of a run whose list of bits to re-evaluate overflowed
    bool ran_fused = false, mags_full = true;   // the last run decided bits in k_emit_chunk; magnitudes of every bit are on the device
    bool ran = false, finished = false;
    int32_t chunk_cap_max = 0;            // largest per-drop iteration capacity (grid of k_emit_chunk)
    bool any_f64in = false;               // a drop takes its samples through axctd_batch_upload_f64 (config decimate = 3)
    bool lent = false;                    // another engine's stream has read this batch's PCM (axctd_batch_copy_from)
    // streaming decode (axctd_batch_stream_*): drops hold growing recordings, `drops` carries their current lengths
    bool streaming = false, stream_closed = false, in_stream_run = false;
    int stream_runs = 0;
    std::vector<int64_t> cap_n;           // samples each drop was created for
    std::vector<int64_t> rows_done;       // rows already on the host
    double* d_norm = nullptr;
    bool ran_nosync = false;              // the last run was enqueued without host round trips: finish() checks the flags
    int64_t n_fallbacks = 0;              // runs that had to be repeated with the host-driven loops
    double ms_total = 0, ms_filter = 0, ms_tone = 0;
    double ms_phase[5] = {0, 0, 0, 0, 0};
#ifndef AXCTD_EMU
    cudaEvent_t ev[6];
    cudaEvent_t evx[2];                   // hand-over to / from the engine's high-priority stream
#endif
};

template <typename T>
static int ax_alloc_arr(axctd_batch* b, T** p, int64_t count) {
    void* v = nullptr;
    if (ax_alloc(b->eng, &v, (size_t)std::max<int64_t>(count, 1) * sizeof(T))) return 1;
    b->allocs.push_back(v);
    *p = (T*)v;
    return 0;
}

// ============================================================ C ABI: engine
extern "C" int axctd_abi_version(void) { return AXCTD_ABI_VERSION; }
extern "C" int axctd_has_cuda(void) {
#ifdef AXCTD_EMU
    return 0;
#else
    return 1;
#endif
}

extern "C" int axctd_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(axctd_config_desc);
        case 1: return (int)sizeof(axctd_drop_summary);
        case 2: return (int)sizeof(axctd_frame);
        case 3: return (int)sizeof(axctd_chunk);
        case 4: return (int)sizeof(axctd_row);
    }
    return -1;
}

extern "C" int axctd_engine_create(int device, axctd_engine** out) {
    if (!out) return AXCTD_ERR_ARG;
    axctd_engine* e = new axctd_engine();
    e->device = device;
#ifndef AXCTD_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        // no silent CPU path: the product needs a CUDA device
        delete e; *out = nullptr; return AXCTD_ERR_CUDA;
    }
    if (ax_fail(e, cudaSetDevice(device), "cudaSetDevice") ||
        ax_fail(e, cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
        delete e; *out = nullptr; return AXCTD_ERR_CUDA;
    }
    {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess || hi >= lo ||
            cudaStreamCreateWithPriority(&e->hp_stream, cudaStreamNonBlocking, hi) != cudaSuccess) { e->hp_stream = nullptr; cudaGetLastError(); }
    }
#endif
    void* p = nullptr;
    if (ax_alloc(e, &p, sizeof(AxCfg) * e->cfg_cap)) { delete e; *out = nullptr; return AXCTD_ERR_CUDA; }
    e->d_cfg = (AxCfg*)p;
#ifndef AXCTD_EMU
    if (device >= 0 && device < 64) { std::lock_guard<std::mutex> g(g_dev_cache[device].mu); g_dev_cache[device].engines++; }
#endif
    *out = e;
    return AXCTD_OK;
}

extern "C" void axctd_engine_destroy(axctd_engine* e) {
    if (!e) return;
    AX_DEV(e);
    for (void* p : e->cfg_allocs) ax_free(e, p);
    ax_free(e, e->d_cfg);
#ifndef AXCTD_EMU
    if (e->stream && e->own_stream) cudaStreamDestroy(e->stream);
    if (e->hp_stream) { cudaStreamSynchronize(e->hp_stream); cudaStreamDestroy(e->hp_stream); }
    if (e->device >= 0 && e->device < 64) {          // the last engine on a device returns the cached blocks
        bool last;
        { std::lock_guard<std::mutex> g(g_dev_cache[e->device].mu); last = --g_dev_cache[e->device].engines <= 0; }
        if (last) ax_pool_trim(e->device);
    }
#endif
    delete e;
}

extern "C" int axctd_engine_set_stream(axctd_engine* e, void* cuda_stream) {
    if (!e) return AXCTD_ERR_ARG;
#ifndef AXCTD_EMU
    if (e->stream && e->own_stream) { cudaStreamSynchronize(e->stream); cudaStreamDestroy(e->stream); }
    e->stream = (cudaStream_t)cuda_stream;
    e->own_stream = false;
#else
    (void)cuda_stream;
#endif
    return AXCTD_OK;
}

extern "C" const char* axctd_last_error(axctd_engine* e) { return e ? e->err.c_str() : "null engine"; }
extern "C" int64_t axctd_engine_launch_count(axctd_engine* e) { return e ? e->launches : 0; }

extern "C" int axctd_engine_set_option(axctd_engine* e, const char* name, double v) {
    if (!e || !name) return AXCTD_ERR_ARG;
    std::string s(name);
    if (s == "segment_len") e->opt_segment_len = (int64_t)v;
    else if (s == "exact_head") e->opt_exact_head = (int)v;
    else if (s == "guard") e->opt_guard = v;
    else if (s == "force_exact") e->opt_force_exact = (int)v;
    else if (s == "tone_direct") e->opt_tone_direct = (int)v;
    else if (s == "max_fixups") e->opt_max_fixups = (int)v;
    else if (s == "filter_variant") e->opt_filter_variant = (int)v;
    else if (s == "zc_div") e->opt_zc_div = std::max(2, (int)v);
    else if (s == "inject_misspec") e->opt_inject_misspec = (int)v;
    else if (s == "ws") e->opt_ws = (int)v;
    else if (s == "fir_first") e->opt_fir_first = (int)v;
    else if (s == "bulk") e->opt_bulk = (int)v;
    else if (s == "heavy_prio") e->opt_heavy_prio = (int)v;
    else if (s == "tone_mma") e->opt_tone_mma = (int)v;
    else if (s == "tone_int8") e->opt_tone_int8 = (int)v;
    else if (s == "tone_complement") e->opt_tone_complement = (int)v;
    else if (s == "seg_target") e->opt_seg_target = (int)v;
    else if (s == "heavy_chain") e->opt_heavy_chain = (int)v;
    else if (s == "pair_launch") e->opt_pair_launch = (int)v;
    else if (s == "scan_only") e->opt_scan_only = (int)v;
    else if (s == "nosync") e->opt_nosync = (int)v;
    else if (s == "fuse_bits") e->opt_fuse_bits = (int)v;
    else if (s == "bit_tol") e->opt_bit_tol = v;
    else if (s == "hist_tol") e->opt_hist_tol = v;
    else if (s == "bitfix_all") e->opt_bitfix_all = (int)v;
    else if (s == "demod_probe") e->opt_demod_probe = (int)v;
#ifndef AXCTD_EMU
    else if (s == "pool") g_pool_on = v != 0.0;                              // block cache on / off (process-wide)
    else if (s == "pool_poison") g_pool_poison = v != 0.0;                   // test hook: recycled blocks are filled with 0xA5
    else if (s == "pool_trim") { AX_DEV(e); cudaStreamSynchronize(e->stream); ax_pool_trim(e->device); }
#else
    else if (s == "pool" || s == "pool_trim" || s == "pool_poison") {}
#endif
    else return AXCTD_ERR_ARG;
    return AXCTD_OK;
}


template <typename T>
static int ax_cfg_upload(axctd_engine* e, const T** dst, const T* src, size_t count) {
    void* p = nullptr;
    if (ax_alloc(e, &p, count * sizeof(T))) return 1;
    e->cfg_allocs.push_back(p);
    if (ax_h2d(e, p, src, count * sizeof(T))) return 1;
    *dst = (const T*)p;
    return 0;
}

extern "C" int axctd_config_create(axctd_engine* e, const axctd_config_desc* ds, int* config_id) {
    if (!e || !ds || !config_id) return AXCTD_ERR_ARG;
    AX_DEV(e);
    if ((int)e->cfgs.size() >= e->cfg_cap) { e->err = "too many configs"; return AXCTD_ERR_CAPACITY; }
    if (ds->n_sections < 1 || ds->n_sections > AX_MAXSEC || ds->bit_inset != 1 || ds->npcm < 1 ||
        ds->bit_cs_len < ds->npcm + 1 || ds->n_power < 1 || ds->d_pcm < 1 || ds->chunk_len < 1 ||
        !ds->bit_cs || !ds->tone_cs || !ds->temp_lut || !ds->hist_edges || !ds->hist_centers ||
        ds->n_hist_edges < 3 || ds->n_hist_edges > 512) { e->err = "bad config"; return AXCTD_ERR_ARG; }
    AxCfg c;
    memset(&c, 0, sizeof(c));
    c.fs = ds->fs;
    c.fs2 = (int64_t)llround(2.0 * ds->fs);
    c.n_power = ds->n_power; c.d_pcm = ds->d_pcm; c.npcm = ds->npcm; c.chunk_len = ds->chunk_len;
    c.pad = ds->pad; c.inset = ds->bit_inset; c.bitrate = ds->bitrate; c.nsec = ds->n_sections;
    for (int s = 0; s < AX_MAXSEC; ++s) for (int q = 0; q < 6; ++q) c.sos[s][q] = (s < c.nsec) ? ds->sos[s][q] : 0.0;
    double r = ds->max_pole_radius;
    if (!(r > 0.0 && r < 1.0)) { e->err = "max_pole_radius must be in (0,1)"; return AXCTD_ERR_ARG; }
    // transient of a zero-state restart relative to the continuous filter decays like r^n:
    // overlap until it is below 1e-19 of full scale (a few hundred times below fp64 epsilon)
    int warm = (int)ceil(log(1e-19) / log(r));
    warm = ((warm + 63) / 64) * 64;
    if (warm < 256) warm = 256;
    c.warm = warm;
    c.head = e->opt_exact_head > 0 ? e->opt_exact_head : warm;
    if (c.head < c.pad + 8) c.head = c.pad + 8;
    c.head_zc_cap = c.head / 4 + 64;
    c.decimate = ds->decimate == 2 ? 2 : ds->decimate == 3 ? 3 : 1;      // 3: the samples arrive as the normalised double-precision signal
    if (c.decimate == 2) {
        if (ds->decim_sections < 1 || ds->decim_sections > AX_MAXSEC || ds->decim_padlen < 1 ||
            !(ds->decim_pole_radius > 0.0 && ds->decim_pole_radius < 1.0)) { e->err = "bad decimator description"; return AXCTD_ERR_ARG; }
        c.dnsec = ds->decim_sections; c.dpad = ds->decim_padlen;
        int dw = (int)ceil(log(1e-19) / log(ds->decim_pole_radius));
        c.dwarm = ((dw + 63) / 64) * 64;
        for (int q = 0; q < c.dnsec; ++q) {
            for (int i = 0; i < 6; ++i) c.dsos[q][i] = ds->decim_sos[q][i];
            c.dzi[q][0] = ds->decim_zi[q][0]; c.dzi[q][1] = ds->decim_zi[q][1];
        }
    }
    AxToneTab ttab;
    memset(&ttab, 0, sizeof(ttab));
    {
        long double ts[6] = {0, 0, 0, 0, 0, 0};
        for (int m = 0; m < c.n_power; ++m)
            for (int q = 0; q < 6; ++q) {
                ts[q] += (long double)ds->tone_cs[6 * (size_t)m + q];
                if (m < AX_TB) ttab.t[m][q] = ds->tone_cs[6 * (size_t)m + q];
            }
        for (int q = 0; q < 6; ++q) c.tone_tsum[q] = (double)ts[q];
        for (int j = 0; j < AX_TONE_ROT; ++j)
            for (int q = 0; q < 6; ++q) c.tone_rot[j][q] = (j * AX_TB < c.n_power) ? ds->tone_cs[6 * (size_t)(j * AX_TB) + q] : 0.0;
    }
    c.min_r400 = ds->min_r400; c.min_dr7500 = ds->min_dr7500;
    c.min_r400_inprof = ds->min_r400 / 2; c.min_dr7500_inprof = ds->min_dr7500 / 2;     // AXCTDprocessor.py:226,228
    c.trig_from = ds->trigger_from_s; c.trig_to = ds->trigger_to_s; c.scale0 = ds->high_bit_scale0;
    for (int q = 0; q < 4; ++q) { c.zc[q] = ds->zcoeff[q]; c.tc[q] = ds->tcoeff[q]; c.cc[q] = ds->ccoeff[q]; }
    for (int q = 0; q < 2; ++q) { c.tlims[q] = ds->tlims[q]; c.slims[q] = ds->slims[q]; }
    const double fs = ds->fs;
    c.off_4p5 = (int64_t)(fs * 4.5); c.off_5p5 = (int64_t)(fs * 5.5);                     // :388-392
    c.off_trig_from = (int64_t)(ds->trigger_from_s * fs); c.off_trig_to = (int64_t)(fs * ds->trigger_to_s);   // :397,404
    c.h1s = (int64_t)(fs * 2.3); c.h1e = (int64_t)(fs * 3.3);                             // :447-448
    c.h2s = (int64_t)(fs * 10.5); c.h2e = (int64_t)(fs * 14.8);                           // :451-452
    c.h3s = (int64_t)(fs * 20); c.h3e = (int64_t)(fs * 24.5);                             // :455-456
    c.half = (int64_t)(fs * 0.5);
    c.lut_len = ds->lut_len; c.n_hist_edges = ds->n_hist_edges;
    if (ds->bit_cs_len < AX_WIN_TAPS) { e->err = "bit_cs table shorter than 48 entries"; return AXCTD_ERR_ARG; }
    {   // ax_window32: phase reference in the middle of the nt aligned taps, tab[k] = (cos, sin)(theta_f (nt/2 - 1/2 - k))
        // for the mark and the space tone, k < nt/2; theta_f from entry 1 of the caller's table (= 2 pi f / f_s)
        const double th1 = atan2(ds->bit_cs[4 + 1], ds->bit_cs[4 + 0]), th2 = atan2(ds->bit_cs[4 + 3], ds->bit_cs[4 + 2]);
        // (a window longer than the table serves scan-only configurations: axctd_batch_run_async refuses to demodulate with it)
        const int nh = ax_win_quads(c.npcm) * 4 <= AX_WIN_TAPS ? ax_win_quads(c.npcm) * 2 : 0;
        for (int k = 0; k < AX_WIN_TAPS; ++k) {
            const double ph = (double)nh - 0.5 - (double)k;
            const bool in = k < nh;
            c.win_tab.t[k].x = in ? (float)cos(th1 * ph) : 0.f; c.win_tab.t[k].y = in ? (float)sin(th1 * ph) : 0.f;
            c.win_tab.t[k].z = in ? (float)cos(th2 * ph) : 0.f; c.win_tab.t[k].w = in ? (float)sin(th2 * ph) : 0.f;
        }
    }
    {   // window responses G_f[d] = sum_m e^{j theta_f m} h[d - (npcm-1-m)] (ax_gwin_*): impulse response h of the
        // cascade in long double, long enough for its tail to fall below 1e-18
        int K = (int)ceil(log(1e-18) / log(r)) + 64;
        K = ((K + 31) / 32) * 32;
        std::vector<long double> zz(2 * AX_MAXSEC, 0.0L), h(K);
        for (int n = 0; n < K; ++n) {
            long double u = (n == 0) ? 1.0L : 0.0L;
            for (int q = 0; q < c.nsec; ++q) {
                const long double b0 = c.sos[q][0], b1 = c.sos[q][1], b2 = c.sos[q][2], a1 = c.sos[q][4], a2 = c.sos[q][5];
                const long double y = b0 * u + zz[2 * q];
                zz[2 * q] = b1 * u - a1 * y + zz[2 * q + 1];
                zz[2 * q + 1] = b2 * u - a2 * y;
                u = y;
            }
            h[n] = u;
        }
        const int GL = K + c.npcm - 1;
        std::vector<double> g(4 * (size_t)GL), gc(4 * (size_t)GL);
        long double run[4] = {0, 0, 0, 0};
        for (int d = 0; d < GL; ++d) {
            long double acc[4] = {0, 0, 0, 0};
            for (int m = 0; m < c.npcm; ++m) {
                const int k = d - (c.npcm - 1 - m);
                if (k < 0 || k >= K) continue;
                const double* t4 = ds->bit_cs + 4 * (size_t)m;
                for (int q = 0; q < 4; ++q) acc[q] += (long double)t4[q] * h[k];
            }
            for (int q = 0; q < 4; ++q) { g[4 * (size_t)d + q] = (double)acc[q]; run[q] += acc[q]; gc[4 * (size_t)d + q] = (double)run[q]; }
        }
        c.g_len = GL;
        if (ax_cfg_upload(e, &c.gtab, g.data(), g.size()) || ax_cfg_upload(e, &c.gcum, gc.data(), gc.size())) return AXCTD_ERR_CUDA;
    }
    {   // python's 10**ex as a float (parse.py:278): int power for ex >= 0, libm pow for ex < 0
        std::vector<double> p10(AX_POW10_LEN);
        for (int ex = -99; ex <= 999; ++ex) {
            if (ex >= 0) { char buf[32]; snprintf(buf, sizeof(buf), "1e%d", ex); p10[ex + 99] = strtod(buf, nullptr); }
            else p10[ex + 99] = pow(10.0, (double)ex);
        }
        if (ax_cfg_upload(e, &c.pow10, p10.data(), p10.size())) return AXCTD_ERR_CUDA;
    }
    if (ax_cfg_upload(e, &c.tone_cs, ds->tone_cs, 6 * (size_t)ds->n_power) ||
        ax_cfg_upload(e, &c.lut, ds->temp_lut, (size_t)ds->lut_len) ||
        ax_cfg_upload(e, &c.hist_edges, ds->hist_edges, (size_t)ds->n_hist_edges) ||
        ax_cfg_upload(e, &c.hist_centers, ds->hist_centers, (size_t)ds->n_hist_edges - 1)) return AXCTD_ERR_CUDA;
    {
        std::vector<double> soa(6 * (size_t)c.n_power);
        for (int m = 0; m < c.n_power; ++m)
            for (int q = 0; q < 6; ++q) soa[(size_t)q * c.n_power + m] = ds->tone_cs[6 * (size_t)m + q];
        if (ax_cfg_upload(e, &c.tone_soa, soa.data(), soa.size())) return AXCTD_ERR_CUDA;
        std::vector<double> t8(8 * (size_t)AX_TB, 0.0);
        for (int m = 0; m < AX_TB && m < c.n_power; ++m)
            for (int q = 0; q < 6; ++q) t8[8 * (size_t)m + q] = ds->tone_cs[6 * (size_t)m + q];
        if (ax_cfg_upload(e, &c.tone_tab8, t8.data(), t8.size())) return AXCTD_ERR_CUDA;
#ifndef AXCTD_EMU
        // k_stats_tones_imma: P = round(p * 2^45) = sum_j d_j 256^j with signed digits; word layout [k-step][digit][lane][2]:
        // lane = 4 g + t holds column g, samples 32 ks + 4 t + i (word 0) and 32 ks + 16 + 4 t + i (word 1) in byte i
        std::vector<uint32_t> ti(AX_STI_TAB_WORDS, 0u);
        for (int m = 0; m < AX_TB && m < c.n_power; ++m)
            for (int q = 0; q < 6; ++q) {
                long long P = llround(ldexp(t8[8 * (size_t)m + q], AX_STI_SHIFT));
                const int ks = m >> 5, kk = m & 31, half = kk >> 4, tt = (kk & 15) >> 2, bi = kk & 3;
                for (int j = 0; j < AX_STI_DIGITS; ++j) {
                    const int dg = (int)(((P & 0xff) ^ 0x80) - 0x80);
                    P = (P - dg) >> 8;
                    ti[(((size_t)ks * AX_STI_DIGITS + j) * 32 + 4 * q + tt) * 2 + half] |= (uint32_t)(dg & 0xff) << (8 * bi);
                }
                if (P != 0) { e->err = "tone phasor outside the digit range"; return AXCTD_ERR_ARG; }
            }
        if (ax_cfg_upload(e, &c.tone_tabi, ti.data(), ti.size())) return AXCTD_ERR_CUDA;
#endif
    }
    e->cfgs.push_back(c);
    e->tone_tabs.push_back(ttab);
    if (ax_h2d(e, e->d_cfg + (e->cfgs.size() - 1), &e->cfgs.back(), sizeof(AxCfg)) || ax_sync(e)) return AXCTD_ERR_CUDA;
    *config_id = (int)e->cfgs.size() - 1;
    return AXCTD_OK;
}

// ============================================================ C ABI: batch
extern "C" void axctd_batch_destroy(axctd_batch* b) {
    if (!b) return;
    AX_DEV(b->eng);
#ifndef AXCTD_EMU
    cudaStreamSynchronize(b->eng->stream);
    if (b->eng->hp_stream) cudaStreamSynchronize(b->eng->hp_stream);
    if (b->lent) cudaDeviceSynchronize();
    for (int i = 0; i < 6; ++i) cudaEventDestroy(b->ev[i]);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(b->evx[i]);
#endif
    for (void* p : b->allocs) ax_free(b->eng, p);
    ax_free(b->eng, b->d_stage);
    ax_host_free(b->h_st); ax_host_free(b->h_row); ax_host_free(b->h_chunk);
    delete b;
}

extern "C" int axctd_batch_create(axctd_engine* e, int n_drops, const int64_t* n_samples, const int32_t* config_id,
                                  axctd_batch** out) {
    if (!e || n_drops <= 0 || !n_samples || !config_id || !out) return AXCTD_ERR_ARG;
    // several kernels put the drop index in gridDim.y (k_stats_tones*, k_tone_*, k_bits_chunk phase 0)
    if (n_drops > 65535) { e->err = "at most 65535 drops per batch"; return AXCTD_ERR_ARG; }
    axctd_batch* b = new axctd_batch();
    b->eng = e; b->n = n_drops;
    memset(&b->w, 0, sizeof(AxWave));
#ifndef AXCTD_EMU
    cudaSetDevice(e->device);
    for (int i = 0; i < 6; ++i) cudaEventCreate(&b->ev[i]);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&b->evx[i], cudaEventDisableTiming);
#endif
    int64_t total = 0;
    int warm_max = 0, head_cap_max = 0, chunk_len_max = 0;
    for (int d = 0; d < n_drops; ++d) {
        if (config_id[d] < 0 || config_id[d] >= (int)e->cfgs.size() || n_samples[d] < 0 || n_samples[d] > 2000000000LL) {
            e->err = "bad drop descriptor"; axctd_batch_destroy(b); return AXCTD_ERR_ARG;
        }
        const AxCfg& c = e->cfgs[config_id[d]];
        total += n_samples[d];
        warm_max = std::max(warm_max, c.warm);
        head_cap_max = std::max(head_cap_max, c.head_zc_cap);
        chunk_len_max = std::max(chunk_len_max, c.chunk_len);
    }
    if (e->opt_force_exact) head_cap_max = chunk_len_max / 4 + 64;
    // segment length of the continuous pass: enough threads to fill the GPU, little warm-up waste
    int64_t L = e->opt_segment_len;
    if (L > 0) L = ((L + 63) / 64) * 64;
    if (L <= 0) {
        // Every lane of the demodulation pass does the same amount of work, so the pass runs in waves of
        // (SMs x 8 warps x 32) segments: take the number of waves that segments of about seg_target samples need and then
        // the shortest segment length whose padded segment count still fits them.  seg_target = 16384 (a 32-drop sub-batch
        // of 12-minute drops runs as two waves of ~14 k-sample segments, warm-up overlap 5 %): measured 14.7 ms per step
        // against 15.3 at 8192 (four waves, overlap 10 %) and 14.3 .. 14.9 for a single wave (runs 42, 43).
        int sms = 148;
#ifndef AXCTD_EMU
        { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && v > 0) sms = v; }
#endif
        const int64_t lanes = (int64_t)sms * 8 * 32;
        auto nseg_for = [&](int64_t len) {
            int64_t s = 0;
            for (int d = 0; d < n_drops; ++d) {
                const int64_t nd = e->cfgs[config_id[d]].decimate == 2 ? (n_samples[d] + 1) / 2 : n_samples[d];
                s += (((nd + len - 1) / len + 127) / 128) * 128;
            }
            return s;
        };
        const int64_t tgt = e->opt_seg_target > 0 ? e->opt_seg_target : 16384;
        const int64_t waves = std::max<int64_t>(1, (total + tgt * lanes - 1) / (tgt * lanes));
        L = 2048;
        while (L < 2 * tgt && nseg_for(L) > waves * lanes) L += 64;
        while (L < 2 * (int64_t)warm_max) L <<= 1;
    }
    AxWave& w = b->w;
    w.n_drops = n_drops; w.n_cfg = (int)e->cfgs.size(); w.cfg = e->d_cfg;
    w.seg_len = (int32_t)L; w.seg_cap = (int32_t)(L / 8 + 32);
    w.guard = e->opt_guard; w.tone_direct = e->opt_tone_direct; w.force_exact = e->opt_force_exact;
    w.bit_tol = e->opt_bit_tol; w.hist_tol = e->opt_hist_tol; w.bitfix_all = e->opt_bitfix_all;
    w.probe = e->opt_demod_probe;
    w.tone_complement = e->opt_tone_complement;
    w.head_zc_cap_max = head_cap_max;
    b->drops.resize(n_drops);
    int64_t pcm_off = 0, zc_off = 0, edge_off = 0, tb_off = 0, xf_off = 0, fwd_off = 0, zq_off = 0;
    int32_t ntb_max = 0, dseg_off = 0;
    // samples per decimation segment: one wave of (SMs x 8 warps x 32) lanes over the recordings that are halved, but
    // not below 2048 (the warm-up overlap is some 650 samples)
    int64_t DL = 8192;
    {
        int64_t dec_total = 0;
        for (int d = 0; d < n_drops; ++d) if (e->cfgs[config_id[d]].decimate == 2) dec_total += n_samples[d];
        int sms = 148;
#ifndef AXCTD_EMU
        { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, e->device) == cudaSuccess && v > 0) sms = v; }
#endif
        const int64_t lanes = (int64_t)sms * 8 * 32;
        DL = ((dec_total / lanes + 1 + 63) / 64) * 64;
        if (DL < 2048) DL = 2048;
        if (DL > 65536) DL = 65536;
    }
    w.dseg_len = (int32_t)DL;
    int32_t seg_off = 0, slab_off = 0, tile_off = 0, chunk_off = 0, pw_off = 0, frame_off = 0;
    for (int d = 0; d < n_drops; ++d) {
        const AxCfg& c = e->cfgs[config_id[d]];
        AxDrop& dr = b->drops[d];
        const int64_t n_raw = n_samples[d];
        const bool dec = c.decimate == 2;
        const bool f64in = c.decimate == 3;                           // axctd_batch_upload_f64: no int16 samples at all
        const int64_t n = dec ? (n_raw + 1) / 2 : n_raw;              // len(y[::2])
        dr.pcm_off = pcm_off; dr.n = n; dr.n_raw = f64in ? 0 : n_raw; dr.cfg = config_id[d];
        pcm_off += f64in ? 64 : ((n_raw + 63) / 64) * 64 + 64;
        dr.xf_off = -1; dr.fwd_off = -1; dr.dseg_base = dseg_off; dr.ndseg = 0;
        if (f64in) { dr.xf_off = xf_off; xf_off += ((n + 63) / 64) * 64 + 64; b->any_f64in = true; }
        if (dec) {
            if (n_raw < 2 * (int64_t)c.dpad + 2) { e->err = "recording too short to decimate"; axctd_batch_destroy(b); return AXCTD_ERR_ARG; }
            const int64_t E = n_raw + 2 * (int64_t)c.dpad;
            dr.xf_off = xf_off; xf_off += ((n + 63) / 64) * 64 + 64;
            dr.fwd_off = fwd_off; fwd_off += ((E + 8 + 1) / 2) * 2;                  // (even: 16-byte rows of the backward pass)
            dr.ndseg = (int32_t)((E + DL - 1) / DL); dseg_off += ((dr.ndseg + 31) / 32) * 32;   // a warp of k_decim_fused stays inside one drop
        }
        dr.seg_base = seg_off; dr.nseg = (int32_t)((n + L - 1) / L); seg_off += ((dr.nseg + 127) / 128) * 128;
        dr.slab_base = slab_off; dr.nslab = (int32_t)(((f64in ? 0 : n_raw) + AX_STAT_SLAB - 1) / AX_STAT_SLAB); slab_off += dr.nslab;
        dr.tb_base = tb_off; dr.ntb = (int32_t)(n / AX_TB); tb_off += dr.ntb; ntb_max = std::max(ntb_max, (int32_t)(((f64in ? 0 : n_raw) + AX_TB - 1) / AX_TB));
        dr.zc_base = zc_off; dr.zc_cap = n / e->opt_zc_div + 4096; zc_off += ((dr.zc_cap + 8 + 63) / 64) * 64;   // (64-aligned: k_tiles_reg loads a tile's walk steps as four 16-byte words)
        dr.zq_base = zq_off; zq_off += n / AX_ZQ + 2;
        dr.tile_base = tile_off; dr.tile_cap = (int32_t)(dr.zc_cap / AX_TILE + 1); tile_off += dr.tile_cap;
        dr.chunk_base = chunk_off; dr.chunk_cap = (int32_t)(2 * (n / c.chunk_len) + 16); chunk_off += dr.chunk_cap;
        b->chunk_cap_max = std::max(b->chunk_cap_max, dr.chunk_cap);
        dr.edge_base = edge_off; dr.edge_cap = n / 24 + 8 * (int64_t)dr.chunk_cap; edge_off += ((dr.edge_cap + 64 + 63) / 64) * 64;
        dr.pw_base = pw_off; dr.pw_cap = (int32_t)(n / c.d_pcm + 2 * (int64_t)dr.chunk_cap + 8); pw_off += dr.pw_cap;
        dr.frame_base = frame_off; dr.frame_cap = (int32_t)(dr.edge_cap / 32 + 64); frame_off += dr.frame_cap;
    }
    b->pcm_total = pcm_off; b->zc_total = zc_off; b->tile_total = tile_off; b->chunk_total = chunk_off;
    b->edge_total = edge_off; b->frame_total = frame_off;
    w.nseg_total = seg_off; w.nslab_total = slab_off; w.pw_total = pw_off; w.ntb_max = ntb_max;
    b->tb_total = tb_off; b->dseg_total = dseg_off;
    std::vector<int32_t> seg_drop(seg_off), slab_drop(slab_off);
    for (int d = 0; d < n_drops; ++d) {
        for (int s = 0; s < ((b->drops[d].nseg + 127) / 128) * 128; ++s) seg_drop[b->drops[d].seg_base + s] = d;
        for (int s = 0; s < b->drops[d].nslab; ++s) slab_drop[b->drops[d].slab_base + s] = d;
    }
    int bad = 0;
    AxDrop* d_drop; int32_t* d_seg_drop; int32_t* d_slab_drop; AxState* d_st;
    bad |= ax_alloc_arr(b, &d_drop, n_drops);
    bad |= ax_alloc_arr(b, &d_st, n_drops);
    bad |= ax_alloc_arr(b, &b->d_pcm, pcm_off + 64);
    bad |= ax_alloc_arr(b, &d_seg_drop, seg_off);
    bad |= ax_alloc_arr(b, &d_slab_drop, slab_off);
    bad |= ax_alloc_arr(b, &w.xf, xf_off + 8);
    bad |= ax_alloc_arr(b, &w.fwd, fwd_off + 8);
    bad |= ax_alloc_arr(b, &w.seg_cnt, seg_off);
    bad |= ax_alloc_arr(b, &w.seg_off, seg_off);
    bad |= ax_alloc_arr(b, &w.blk_sum, seg_off / 128 + 1);
    bad |= ax_alloc_arr(b, &w.seg_unc, seg_off);
    bad |= ax_alloc_arr(b, &w.head_unc, chunk_off);
    bad |= ax_alloc_arr(b, &w.unc_list, 2 * (int64_t)AX_UNC_CAP * n_drops);
    const int64_t rec_total = (int64_t)seg_off * w.seg_cap;
    bad |= ax_alloc_arr(b, &w.rec_idx, rec_total);
    bad |= ax_alloc_arr(b, &w.rec_a1, rec_total);
    bad |= ax_alloc_arr(b, &w.rec_a2, rec_total);
    bad |= ax_alloc_arr(b, &w.zc_idx, zc_off);
    bad |= ax_alloc_arr(b, &w.zc_a1, zc_off);
    bad |= ax_alloc_arr(b, &w.zc_a2, zc_off);
    bad |= ax_alloc_arr(b, &w.zc_nx, zc_off + 64);
    bad |= ax_alloc_arr(b, &w.zc_q, zq_off + 8);
    bad |= ax_alloc_arr(b, &w.tile_mask, (int64_t)tile_off * 4);
    bad |= ax_alloc_arr(b, &w.tile_map, tile_off);
    bad |= ax_alloc_arr(b, &w.cmask, (int64_t)tile_off + 8);
    bad |= ax_alloc_arr(b, &w.crank, (int64_t)tile_off + 8);
    bad |= ax_alloc_arr(b, &w.chunk, chunk_off);
    bad |= ax_alloc_arr(b, &w.head_idx, (int64_t)chunk_off * head_cap_max);
    bad |= ax_alloc_arr(b, &w.head_a1, (int64_t)chunk_off * head_cap_max);
    bad |= ax_alloc_arr(b, &w.head_a2, (int64_t)chunk_off * head_cap_max);
    bad |= ax_alloc_arr(b, &w.head_cnt, chunk_off);
    bad |= ax_alloc_arr(b, &w.pw_raw, 3 * (int64_t)pw_off);
    bad |= ax_alloc_arr(b, &w.pw_sm, 3 * (int64_t)pw_off);
    bad |= ax_alloc_arr(b, &w.r400, pw_off);
    bad |= ax_alloc_arr(b, &w.r7500, pw_off);
    bad |= ax_alloc_arr(b, &w.pw_ind, pw_off);
    bad |= ax_alloc_arr(b, &w.tone_rng, 2 * (int64_t)n_drops);
    bad |= ax_alloc_arr(b, &w.tone_acc, 6 * (int64_t)pw_off);
    bad |= ax_alloc_arr(b, &w.tb_sum, tb_off * 6 + 8);
    bad |= ax_alloc_arr(b, &w.edge_idx, edge_off);
    bad |= ax_alloc_arr(b, &w.lvl_slot, edge_off);
    bad |= ax_alloc_arr(b, &w.r7500m, pw_off);
    bad |= ax_alloc_arr(b, &w.bit, edge_off);
    bad |= ax_alloc_arr(b, &w.a1, edge_off);
    bad |= ax_alloc_arr(b, &w.a2, edge_off);
    bad |= ax_alloc_arr(b, &w.bitw, edge_off / 32 + 4);
    bad |= ax_alloc_arr(b, &w.validw, edge_off / 32 + 4);
    bad |= ax_alloc_arr(b, &w.frame, frame_off);
    bad |= ax_alloc_arr(b, &w.row, frame_off);
    bad |= ax_alloc_arr(b, &w.chunk_out, chunk_off);
    b->h_st = (AxState*)ax_host_alloc(sizeof(AxState) * (size_t)n_drops);
    b->h_row = (axctd_row*)ax_host_alloc(sizeof(axctd_row) * (size_t)frame_off);
    b->h_chunk = (axctd_chunk*)ax_host_alloc(sizeof(axctd_chunk) * (size_t)chunk_off);
    if (!b->h_st || !b->h_row || !b->h_chunk) bad = 1;
    bad |= ax_alloc_arr(b, &b->d_qc, 2 * (int64_t)frame_off + 16);
    bad |= ax_alloc_arr(b, &w.flags, 8);
    w.fix_cap = edge_off / 8 + 1024;
    bad |= ax_alloc_arr(b, &w.fix_list, w.fix_cap);
    if (bad) { axctd_batch_destroy(b); return AXCTD_ERR_CUDA; }
    w.drop = d_drop; w.st = d_st; w.pcm = b->d_pcm; w.seg_drop = d_seg_drop; w.slab_drop = d_slab_drop;
    bad |= ax_h2d(e, d_drop, b->drops.data(), sizeof(AxDrop) * n_drops);
    bad |= ax_h2d(e, d_seg_drop, seg_drop.data(), sizeof(int32_t) * seg_off);
    bad |= ax_h2d(e, d_slab_drop, slab_drop.data(), sizeof(int32_t) * slab_off);
    bad |= ax_zero(e, b->d_pcm, sizeof(int16_t) * (pcm_off + 64));
    bad |= ax_sync(e);
    if (bad) { axctd_batch_destroy(b); return AXCTD_ERR_CUDA; }
    b->st.resize(n_drops);
    b->summary.resize(n_drops);
    *out = b;
    return AXCTD_OK;
}

extern "C" int axctd_batch_upload(axctd_batch* b, int drop, const int16_t* pcm, int64_t n) {
    if (!b || b->streaming || drop < 0 || drop >= b->n || !pcm || n != b->drops[drop].n_raw ||
        b->eng->cfgs[b->drops[drop].cfg].decimate == 3) return AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    if (ax_h2d(b->eng, b->d_pcm + b->drops[drop].pcm_off, pcm, sizeof(int16_t) * n)) return AXCTD_ERR_CUDA;
    b->ran = false;
    return AXCTD_OK;
}

extern "C" int axctd_batch_upload_f64(axctd_batch* b, int drop, const double* samples, int64_t n) {
    if (!b || b->streaming || drop < 0 || drop >= b->n || !samples || b->eng->cfgs[b->drops[drop].cfg].decimate != 3 ||
        n != b->drops[drop].n) return AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    if (ax_h2d(b->eng, b->w.xf + b->drops[drop].xf_off, samples, sizeof(double) * n)) return AXCTD_ERR_CUDA;
    b->ran = false;
    return AXCTD_OK;
}

extern "C" int axctd_batch_upload_interleaved(axctd_batch* b, int drop, const int16_t* frames, int64_t n_frames, int channels) {
    if (!b || b->streaming || drop < 0 || drop >= b->n || !frames || channels < 1 || n_frames != b->drops[drop].n_raw ||
        b->eng->cfgs[b->drops[drop].cfg].decimate == 3) return AXCTD_ERR_ARG;
    if (channels == 1) return axctd_batch_upload(b, drop, frames, n_frames);
    axctd_engine* e = b->eng;
    AX_DEV(e);
    const size_t bytes = sizeof(int16_t) * (size_t)n_frames * channels;
    if (bytes > b->stage_bytes) {                          // staging area for the interleaved frames, grown on demand
        if (ax_sync(e)) return AXCTD_ERR_CUDA;             // (an earlier de-interleave may still be reading the old one)
        ax_free(e, b->d_stage); b->d_stage = nullptr; b->stage_bytes = 0;
        if (ax_alloc(e, &b->d_stage, bytes + 64)) return AXCTD_ERR_CUDA;
        b->stage_bytes = bytes;
    }
    if (ax_h2d(e, b->d_stage, frames, bytes)) return AXCTD_ERR_CUDA;
    const int64_t launches_before = e->launches;
    AX_LAUNCH(e, k_deinterleave, (n_frames + 7) / 8, (const int16_t*)b->d_stage, b->d_pcm + b->drops[drop].pcm_off, n_frames, channels);
    e->launches = launches_before + 1;
    if (ax_launch_check(e)) return AXCTD_ERR_CUDA;
    b->ran = false;
    return AXCTD_OK;
}

extern "C" int axctd_batch_copy_from(axctd_batch* b, int drop, axctd_batch* src, int src_drop, int64_t src_offset, int64_t n) {
    if (!b || !src || b->streaming || drop < 0 || drop >= b->n || src_drop < 0 || src_drop >= src->n || n != b->drops[drop].n_raw ||
        src_offset < 0 || src_offset + n > src->drops[src_drop].n_raw || b->eng->device != src->eng->device ||
        b->eng->cfgs[b->drops[drop].cfg].decimate == 3) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    AX_DEV(e);
    const int16_t* from = src->d_pcm + src->drops[src_drop].pcm_off + src_offset;
    int16_t* to = b->d_pcm + b->drops[drop].pcm_off;
#ifndef AXCTD_EMU
    if (src->eng->stream != e->stream) {                      // the source's pending upload / de-interleave must have landed
        src->lent = true;                                     // (and its blocks must not be recycled under this copy)
        cudaEvent_t ev;
        if (ax_fail(e, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "event") ) return AXCTD_ERR_CUDA;
        cudaEventRecord(ev, src->eng->stream); cudaStreamWaitEvent(e->stream, ev, 0); cudaEventDestroy(ev);
    }
    if (ax_fail(e, cudaMemcpyAsync(to, from, sizeof(int16_t) * (size_t)n, cudaMemcpyDeviceToDevice, e->stream), "D2D")) return AXCTD_ERR_CUDA;
#else
    memcpy(to, from, sizeof(int16_t) * (size_t)n);
#endif
    b->ran = false;
    return AXCTD_OK;
}

extern "C" int axctd_batch_device_pcm(axctd_batch* b, int drop, void** dptr) {
    if (!b || drop < 0 || drop >= b->n || !dptr) return AXCTD_ERR_ARG;
    *dptr = (void*)(b->d_pcm + b->drops[drop].pcm_off);
    return AXCTD_OK;
}

#ifndef AXCTD_EMU
#define AX_EVENT(b, i) cudaEventRecord((b)->ev[i], (b)->eng->stream)
// Several engines (streams) of one process may decode sub-batches concurrently (batch.ConcurrentDecoder): their
// small latency-bound kernels and result downloads overlap each other's big kernels.  The demodulation pass fills
// the GPU by itself, so two of them side by side only stretch each other; every engine therefore waits for the
// pass most recently enqueued by another engine on the same device before it starts its own.
#include <mutex>
static std::mutex g_heavy_mu;
static cudaEvent_t g_heavy_ev[64];
static bool g_heavy_have[64];
static const axctd_engine* g_heavy_owner[64];
static void ax_heavy_begin(axctd_engine* e) {
    if (!e->opt_heavy_chain || e->device < 0 || e->device >= 64) return;
    g_heavy_mu.lock();
    if (g_heavy_have[e->device] && g_heavy_owner[e->device] != e) cudaStreamWaitEvent(e->stream, g_heavy_ev[e->device], 0);
}
static void ax_heavy_end(axctd_engine* e) {
    if (!e->opt_heavy_chain || e->device < 0 || e->device >= 64) return;
    if (!g_heavy_have[e->device]) { cudaEventCreateWithFlags(&g_heavy_ev[e->device], cudaEventDisableTiming); g_heavy_have[e->device] = true; }
    cudaEventRecord(g_heavy_ev[e->device], e->stream);
    g_heavy_owner[e->device] = e;
    g_heavy_mu.unlock();
}
#else
#define ax_heavy_begin(e) ((void)0)
#define ax_heavy_end(e) ((void)0)
#define AX_EVENT(b, i) ((void)0)
#endif

static int ax_run_tones(axctd_batch* b, int phase_b) {
    axctd_engine* e = b->eng;
    AxWave& w = b->w;
    if (!e->opt_tone_direct) {
        bool all_blocked = true;
        for (const AxDrop& dr : b->drops) if (!ax_tone_blocked_ok(e->cfgs[dr.cfg])) all_blocked = false;
#ifndef AXCTD_EMU
        {   // per-drop power-sample range this launch can touch: detection rounds only cover chunks [pa_lo, pa_hi) of the fixed grid
            int i_lo = 0, i_hi = 0;
            for (const AxDrop& dr : b->drops) {
                const AxCfg& c = e->cfgs[dr.cfg];
                const int per = c.chunk_len / c.d_pcm + 2;
                i_hi = std::max(i_hi, phase_b ? dr.pw_cap : (int)std::min<int64_t>((int64_t)w.pa_hi * per, dr.pw_cap));
            }
            if (!phase_b) {
                i_lo = 0x7fffffff;
                for (const AxDrop& dr : b->drops) {
                    const AxCfg& c = e->cfgs[dr.cfg];
                    const int span = c.chunk_len - c.n_power;
                    const int per_min = span > 0 ? span / c.d_pcm : 0;          // a full fixed-grid chunk holds at least this many
                    i_lo = std::min(i_lo, (int)std::min<int64_t>((int64_t)w.pa_lo * per_min, dr.pw_cap));
                }
            }
            if (i_hi > i_lo) {
                k_tone_range<<<(b->n + 127) / 128, 128, 0, e->stream>>>(w, phase_b);
                if (e->opt_tone_mma)
                    k_tone_windows_mma<<<dim3((unsigned)((i_hi - i_lo + 31) / 32), (unsigned)b->n), 128, 0, e->stream>>>(w, i_lo, i_hi);
                else
                    k_tone_windows<<<dim3((unsigned)(((int64_t)(i_hi - i_lo) * 32 + 255) / 256), (unsigned)b->n), 256, 0, e->stream>>>(w, i_lo, i_hi);
                k_tone_mag<<<dim3((unsigned)((i_hi - i_lo + 127) / 128), (unsigned)b->n), 128, 0, e->stream>>>(w, i_lo, i_hi);
                e->launches += 3;
            }
        }
#else
        AX_LAUNCH(e, k_tonewin, (int64_t)w.pw_total, w, phase_b);
#endif
        if (all_blocked) return 0;
        AX_LAUNCH(e, k_tone_direct, (int64_t)w.pw_total, w, phase_b + 2);   // only configs the blocked path skipped
        return 0;
    }
    AX_LAUNCH(e, k_tone_direct, (int64_t)w.pw_total, w, phase_b);
    return 0;
}

extern "C" int axctd_batch_run_async(axctd_batch* b) {
    if (!b) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    AxWave& w = b->w;
    const int n = b->n;
#ifndef AXCTD_EMU
    cudaSetDevice(e->device);
#endif
    b->finished = false;
    const bool streaming = b->streaming;
    if (streaming && !b->in_stream_run) { e->err = "a streaming batch runs through axctd_batch_stream_run"; return AXCTD_ERR_STATE; }
    AX_EVENT(b, 0);
    if (ax_zero(e, w.flags, sizeof(int32_t) * 8)) return AXCTD_ERR_CUDA;
#ifndef AXCTD_EMU
#define AX_INIT() do { k_init_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } while (0)
#else
#define AX_INIT() AX_LAUNCH(e, k_init, n, w)
#endif
    if (!streaming) { AX_INIT(); }
    else if (b->stream_runs == 0) { AX_INIT(); AX_LAUNCH(e, k_stream_begin, n, w, (const double*)b->d_norm); }
    else { AX_LAUNCH(e, k_stream_resume, n, w); }
#ifndef AXCTD_EMU
    {   // one pass over the PCM: statistics and the tone block sums, one launch per rate class in use
        for (size_t ci = 0; ci < e->cfgs.size(); ++ci) {
            if (!std::any_of(b->drops.begin(), b->drops.end(), [&](const AxDrop& d) { return d.cfg == (int)ci; })) continue;
            if (e->cfgs[ci].decimate == 3 || w.ntb_max <= 0) continue;      // drops given as a double-precision signal hold no int16 samples
            const dim3 stm_grid((unsigned)((w.ntb_max + AX_STM_GROUPS * AX_ST_THREADS - 1) / (AX_STM_GROUPS * AX_ST_THREADS)), (unsigned)n);
#define AX_HYB(KD) do { ax_optin_smem<k_stats_tones_hyb<KD>>(AX_STM_SMEM, e->device); \
                        k_stats_tones_hyb<KD><<<stm_grid, AX_ST_THREADS, AX_STM_SMEM, e->stream>>>(w, e->cfgs[ci].tone_tab8, e->tone_tabs[ci], (int)ci); } while (0)
            if (e->opt_tone_mma >= 2) {           // block sums split between the tensor cores and the vector pipe: 4 * (value) of every 64 samples on the tensor cores
                if (e->opt_tone_mma <= 8) AX_HYB(8); else if (e->opt_tone_mma <= 10) AX_HYB(10); else if (e->opt_tone_mma <= 12) AX_HYB(12); else AX_HYB(14);
            } else if (e->opt_tone_mma && e->opt_tone_int8) {
                ax_optin_smem<k_stats_tones_imma>(AX_STI_SMEM, e->device);
                const dim3 sti_grid((unsigned)((w.ntb_max + AX_STI_GROUPS * AX_ST_THREADS - 1) / (AX_STI_GROUPS * AX_ST_THREADS)), (unsigned)n);
                k_stats_tones_imma<<<sti_grid, AX_ST_THREADS, AX_STI_SMEM, e->stream>>>(w, e->cfgs[ci].tone_tabi, (int)ci);
            } else if (e->opt_tone_mma) {
                ax_optin_smem<k_stats_tones_mma>(AX_STM_SMEM, e->device);
                k_stats_tones_mma<<<stm_grid, AX_ST_THREADS, AX_STM_SMEM, e->stream>>>(w, e->cfgs[ci].tone_tab8, (int)ci);
            } else
                k_stats_tones<<<dim3((unsigned)((w.ntb_max + AX_ST_THREADS - 1) / AX_ST_THREADS), (unsigned)n), AX_ST_THREADS, 0, e->stream>>>(w, e->tone_tabs[ci], (int)ci);
            e->launches++;
        }
        if (w.nslab_total > 0 && !streaming) { k_stats_wrap<<<dim3(AX_WRAP_CTAS, (unsigned)n), 256, 0, e->stream>>>(w); e->launches++; }
    }
#else
    if (!streaming) { AX_LAUNCH(e, k_stats, (int64_t)w.nslab_total, w); }
#endif
    if (!streaming) { AX_LAUNCH(e, k_stats_fin, n, w); }        // (streaming: the normalisation was fixed by the caller)
    const bool any_dec = b->dseg_total > 0 || b->any_f64in;      // drops whose samples are doubles (w.xf)
    if (b->dseg_total > 0) {       // recordings above 50 kHz: halve them on the device (AXCTDprocessor.py:60-62)
#ifndef AXCTD_EMU
        bool dfused = e->opt_filter_variant == 0;
        int dpar = -1;
        for (const AxDrop& dr : b->drops) if (dr.xf_off >= 0) {
            const AxCfg& c = e->cfgs[dr.cfg];
            if (!ax_decim_fused_ok(c) || (dpar >= 0 && dpar != (c.dpad & 1))) dfused = false;
            dpar = c.dpad & 1;
        }
        if (dfused) {
            ax_launch_decim_fused<0>(w, b->dseg_total, dpar, e->stream, e->device);
            ax_launch_decim_fused<1>(w, b->dseg_total, dpar, e->stream, e->device);
            e->launches += 2;
        } else
#endif
        {
            AX_LAUNCH(e, k_decim, b->dseg_total, w, 0);
            AX_LAUNCH(e, k_decim, b->dseg_total, w, 1);
        }
    }
    if (any_dec) {
        AX_LAUNCH(e, k_decim_fin, n, w);
#ifndef AXCTD_EMU
        w.only_xf = 1;
        AX_LAUNCH(e, k_toneblock, b->tb_total, w);
        w.only_xf = 0;
#endif
    }
#ifdef AXCTD_EMU
    AX_LAUNCH(e, k_toneblock, b->tb_total, w);
#endif
#ifndef AXCTD_EMU
    // The demodulation pass runs on the engine's high-priority stream: with several sub-batches in flight its CTAs get
    // the SMs ahead of the other engines' kernels, which fill in behind it.
    cudaStream_t main_stream = e->stream;
    if (e->hp_stream && e->opt_heavy_prio) {
        cudaEventRecord(b->evx[0], main_stream); cudaStreamWaitEvent(e->hp_stream, b->evx[0], 0);
        e->stream = e->hp_stream;
    }
#define AX_HEAVY_LEAVE() do { if (e->stream != main_stream) { cudaEventRecord(b->evx[1], e->stream); e->stream = main_stream; \
                                                              cudaStreamWaitEvent(main_stream, b->evx[1], 0); } } while (0)
#else
#define AX_HEAVY_LEAVE() ((void)0)
#endif
    ax_heavy_begin(e);
    AX_EVENT(b, 1);
    const bool scan_only = e->opt_scan_only != 0;     // tone levels only: no crossings are produced
    if (!scan_only)
        for (const AxDrop& dr : b->drops)
            if (ax_win_quads(e->cfgs[dr.cfg].npcm) * 4 > AX_WIN_TAPS) {
                e->err = "bit window longer than the window table (npcm > 45): this rate can only be scanned, not demodulated";
                ax_heavy_end(e); AX_HEAVY_LEAVE();
                return AXCTD_ERR_ARG;
            }
    if (scan_only) { if (ax_zero(e, w.seg_cnt, sizeof(int32_t) * (size_t)w.nseg_total)) { ax_heavy_end(e); AX_HEAVY_LEAVE(); return AXCTD_ERR_CUDA; } }
#ifndef AXCTD_EMU
    bool fused = e->opt_filter_variant == 0;
    std::vector<int> used_cfg;
    for (size_t ci = 0; ci < e->cfgs.size(); ++ci)
        if (std::any_of(b->drops.begin(), b->drops.end(), [&](const AxDrop& d) { return d.cfg == (int)ci; })) {
            used_cfg.push_back((int)ci);
            if (!ax_demod_fused_ok(e->cfgs[ci])) fused = false;
        }
    if (scan_only) {
    } else if (fused) {
        // two low-pass rate classes with window lengths 39 and 43 (a batch of 44.1 and 48 kHz drops): one launch for both
        int pair39 = -1, pair43 = -1;
        if (e->opt_pair_launch && !e->opt_ws && !e->opt_bulk && e->opt_fir_first && used_cfg.size() == 2) {
            const AxCfg& ca = e->cfgs[used_cfg[0]]; const AxCfg& cb = e->cfgs[used_cfg[1]];
            const bool i16_only = std::none_of(b->drops.begin(), b->drops.end(), [&](const AxDrop& d) { return d.xf_off >= 0; });
            if (i16_only && ax_demod_fast_ok(ca) && ax_demod_fast_ok(cb) && ca.npcm + cb.npcm == 82 && (ca.npcm == 39 || ca.npcm == 43)) {
                pair39 = ca.npcm == 39 ? used_cfg[0] : used_cfg[1];
                pair43 = ca.npcm == 39 ? used_cfg[1] : used_cfg[0];
            }
        }
        if (pair39 >= 0) {
            ax_launch_demod_fused_pair<3, true>(w, e->cfgs[pair39], pair39, e->cfgs[pair43], pair43, e->stream, e->device);
            e->launches++;
        } else
        // one launch per rate class in use (CTAs of the other classes exit at once)
        for (int ci : used_cfg) {
            const bool has_i16 = std::any_of(b->drops.begin(), b->drops.end(), [&](const AxDrop& d) { return d.cfg == ci && d.xf_off < 0; });
            const bool has_f64 = std::any_of(b->drops.begin(), b->drops.end(), [&](const AxDrop& d) { return d.cfg == ci && d.xf_off >= 0; });
            if (has_i16) { ax_launch_demod_fused_any<false>(w, e->cfgs[ci], ci, 0, e->stream, e->device, e->opt_ws, e->opt_fir_first, e->opt_bulk); e->launches++; }
            if (has_f64) { ax_launch_demod_fused_f64(w, e->cfgs[ci], ci, e->stream, e->device); e->launches++; }      // halved recordings
        }
    } else
#else
    if (!scan_only)
#endif
    { AX_LAUNCH(e, k_filter, (int64_t)w.nseg_total, w); }
    AX_EVENT(b, 2);
    ax_heavy_end(e);
    AX_HEAVY_LEAVE();
    AX_LAUNCH(e, k_scan_block, (int64_t)w.nseg_total / 128, w);
    AX_LAUNCH1(e, k_scan, n, w);
#ifndef AXCTD_EMU
    k_compact_warp<<<(unsigned)(((int64_t)w.nseg_total * 32 + 255) / 256), 256, 0, e->stream>>>(w);     // (+ walk steps)
    e->launches += 1;
#else
    AX_LAUNCH(e, k_compact, (int64_t)w.nseg_total, w);
    AX_LAUNCH(e, k_nx, b->zc_total, w);
#endif
#ifndef AXCTD_EMU
    if (b->tile_total > 0) { k_tiles_reg<<<(unsigned)((b->tile_total + 127) / 128), 128, 0, e->stream>>>(b->tile_total, w); e->launches++; }
#else
    AX_LAUNCH(e, k_tiles, b->tile_total, w);
#endif
#ifndef AXCTD_EMU
    k_plan0_block<<<n, 128, 0, e->stream>>>(w); e->launches++;
#else
    AX_LAUNCH1(e, k_plan0, n, w);
#endif
    AX_LAUNCH(e, k_pwfill, b->chunk_total, w, 0);
    AX_EVENT(b, 3);
    int32_t flags[8];
    // Two data-dependent loops steer a decode: the 400 Hz pulse search (rounds of fixed-grid chunks until every drop has
    // found its pulse) and the chunk-chain repair (predict / recompute heads / verify until no prediction was wrong).
    // Host-driven, each iteration ends with a flag read.  Without round trips (nosync, the default) a fixed schedule is
    // enqueued instead -- two search rounds (the first 32 iterations of the fixed grid: 64 s at the default chunk) and
    // two chain iterations (one repair), whose kernels do nothing for drops that are already settled -- and
    // axctd_batch_finish looks at the flags once: a drop that needed more repeats the run with the host-driven loops.
    const bool nosync = e->opt_nosync != 0 && !b->force_sync && !streaming;
    b->ran_nosync = nosync;
    // 400 Hz pulse search on the fixed chunk grid, in rounds of chunks (most drops need one round); a streaming run looks
    // at the iterations that are new since the previous run, in one round
    int lo0 = 0, hi0 = 8;
    if (streaming) {
        hi0 = 1 << 28; lo0 = hi0;
        for (int d2 = 0; d2 < n; ++d2) if (b->stream_runs == 0 || b->st[d2].sm_status == 0) lo0 = std::min(lo0, b->stream_runs == 0 ? 0 : (int)b->st[d2].next_sm_chunk);
    }
    for (int lo = lo0, hi = hi0, round = 0; lo < hi; lo = hi, hi = hi * 4, ++round) {
        w.pa_lo = lo; w.pa_hi = hi;
        if (nosync && round > 0 && ax_zero(e, w.flags + AX_FLAG_MORE, sizeof(int32_t))) return AXCTD_ERR_CUDA;
        ax_run_tones(b, 0);
        AX_LAUNCH(e, k_levels, (int64_t)w.pw_total, w, 0);
#ifndef AXCTD_EMU
        k_sm_search_warp<<<n, 32, 0, e->stream>>>(w); e->launches++;
#endif
        AX_LAUNCH1(e, k_sm, n, w, 0);
        if (nosync) { if (round >= 1) break; continue; }
        if (ax_d2h(e, flags, w.flags, sizeof(flags)) || ax_sync(e)) return AXCTD_ERR_CUDA;
        if (!flags[AX_FLAG_MORE]) break;
        flags[AX_FLAG_MORE] = 0;
        if (ax_h2d(e, w.flags, flags, sizeof(flags))) return AXCTD_ERR_CUDA;
        if (hi > (1 << 28)) break;
    }
    AX_EVENT(b, 4);
    // chunk chain: canonical walk tables, then predict, recompute heads exactly, verify; repeat while repairs happen
#ifndef AXCTD_EMU
    k_canon_block<<<n, AX_CANON_THREADS, 0, e->stream>>>(w); e->launches++;
#else
    AX_LAUNCH1(e, k_canon, n, w);
#endif
    for (int it = 0;; ++it) {
        if (nosync && it > 0 && ax_zero(e, w.flags + AX_FLAG_DIRTY, sizeof(int32_t))) return AXCTD_ERR_CUDA;
#ifndef AXCTD_EMU
        if (e->opt_filter_variant == 0) { k_chain_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } else
#endif
        { AX_LAUNCH1(e, k_chain, n, w); }
        if (e->opt_inject_misspec && it == 0) AX_LAUNCH(e, k_inject, n, w);
#ifndef AXCTD_EMU
        if (e->opt_filter_variant == 0) {
            // heads of the rate classes the fused kernel is instantiated for; the generic form takes the rest
            bool rest = any_dec;
            for (int ci : used_cfg) {
                if (ax_demod_fused_ok(e->cfgs[ci])) { ax_launch_demod_fused_any<true>(w, e->cfgs[ci], ci, b->chunk_total, e->stream, e->device, e->opt_ws, 0, 0); e->launches++; }
                else rest = true;
            }
            if (rest) AX_LAUNCH(e, k_headfilt, b->chunk_total, w, 1);
        } else
#endif
        { AX_LAUNCH(e, k_headfilt, b->chunk_total, w, 0); }
        AX_LAUNCH(e, k_headwalk, b->chunk_total, w);
#ifndef AXCTD_EMU
        if (e->opt_filter_variant == 0) { k_verify_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } else
#endif
        { AX_LAUNCH1(e, k_verify, n, w); }
        if (nosync) { if (it >= 1) break; continue; }
        if (ax_d2h(e, flags, w.flags, sizeof(flags)) || ax_sync(e)) return AXCTD_ERR_CUDA;
        if (!flags[AX_FLAG_DIRTY]) break;
        if (it >= e->opt_max_fixups) { e->err = "chunk chain did not converge"; return AXCTD_ERR_STATE; }
        if (ax_zero(e, w.flags, sizeof(int32_t))) return AXCTD_ERR_CUDA;
    }
    // guard-band samples of the filter passes, now that the iterations are known (almost always none)
    AX_LAUNCH(e, k_unc_resolve, (int64_t)n * AX_UNC_CAP, w);
    AX_LAUNCH(e, k_unc_fin, n, w);
#ifndef AXCTD_EMU
    if (e->opt_filter_variant == 0) { k_plan_tones_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } else
#endif
    { AX_LAUNCH1(e, k_plan_tones, n, w); }
    AX_LAUNCH(e, k_pwfill, b->chunk_total, w, 1);
    ax_run_tones(b, 1);
    AX_LAUNCH(e, k_levels, (int64_t)w.pw_total, w, 1);
    AX_LAUNCH1(e, k_sm, n, w, 1);
#ifndef AXCTD_EMU
    if (e->opt_filter_variant == 0) { k_offsets_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } else
#endif
    { AX_LAUNCH1(e, k_offsets, n, w); }
#ifndef AXCTD_EMU
    {
        // The scale calibration reads the mark / space magnitudes of the iterations that overlap [first pulse + 1.8 s,
        // + 3.8 s] (+ margins): the first nk after k0.  Those keep the two-step form (edges + magnitudes, phase-0
        // re-evaluation, calibration, decisions); every later iteration has its bits decided while its edges are
        // emitted (k_emit_chunk mode 2) and stores magnitudes only for the bits listed for a double-precision window
        // (k_bits_recheck) -- 16 B per bit less written and read back.  A caller that asks for the magnitudes
        // (axctd_batch_bits: conf) has them materialised then (ax_materialise_magnitudes).
        int nk = 1;
        for (int d2 = 0; d2 < n; ++d2) {
            const AxCfg& c2 = e->cfgs[b->drops[d2].cfg];
            const double span = (double)c2.h1e + (double)c2.half + 128.0 * c2.fs / c2.bitrate + (double)c2.chunk_len;
            nk = std::max(nk, (int)(span / (0.9 * (double)c2.chunk_len)) + 3);
        }
        const bool fuse = e->opt_fuse_bits != 0 && !streaming && !w.bitfix_all && !b->force_nofuse;
        b->ran_fused = fuse; b->mags_full = !fuse;
        w.nk_full = nk;
        const dim3 region_a((unsigned)nk, (unsigned)n);
        if (fuse) k_emit_chunk<<<region_a, 128, 0, e->stream>>>(w, 1);            // only the nk iterations after the first demodulated one
        else k_emit_chunk<<<dim3((unsigned)b->chunk_cap_max, (unsigned)b->n), 128, 0, e->stream>>>(w, 0);
        e->launches++;
        k_bits_chunk<<<region_a, 128, 0, e->stream>>>(w, 0, 1); e->launches++;
        k_scale_block<<<n, 128, 0, e->stream>>>(w); e->launches++;
        if (fuse) {
            k_bits_chunk<<<region_a, 128, 0, e->stream>>>(w, 1, 1);
            k_emit_chunk<<<dim3((unsigned)b->chunk_cap_max, (unsigned)b->n), 128, 0, e->stream>>>(w, 2);
            k_bits_recheck<<<592, 128, 0, e->stream>>>(w);
            e->launches += 3;
        } else { k_bits_chunk<<<(unsigned)b->chunk_total, 128, 0, e->stream>>>(w, 1, 0); e->launches++; }
    }
#else
    AX_LAUNCH(e, k_emit, b->chunk_total, w);
    AX_LAUNCH(e, k_bits, b->edge_total, w, 0);
    AX_LAUNCH1(e, k_scale, n, w);
    AX_LAUNCH(e, k_bits, b->edge_total, w, 1);
#endif
#ifndef AXCTD_EMU
    k_headers_warp<<<2 * n, 32, 0, e->stream>>>(w); e->launches++;
#else
    AX_LAUNCH1(e, k_headers, 2 * (int64_t)n, w);
#endif
    AX_LAUNCH(e, k_merge, n, w);          // header text -> calibration coefficients (python float semantics, ax_merge_item)
    AX_LAUNCH(e, k_pack, b->edge_total / 32, w);
    AX_LAUNCH(e, k_valid, b->edge_total / 32, w);
    AX_LAUNCH(e, k_frames_spec, b->chunk_total, w);
#ifndef AXCTD_EMU
    if (e->opt_filter_variant == 0) { k_frames_chain_warp<<<n, 32, 0, e->stream>>>(w); e->launches++; } else
#endif
    { AX_LAUNCH1(e, k_frames_chain, n, w); }
    AX_LAUNCH(e, k_frames_write, b->chunk_total, w);
    AX_LAUNCH(e, k_calib, b->frame_total, w);
#ifndef AXCTD_EMU
    if (e->opt_filter_variant == 0) { k_qc_warp<<<(unsigned)((b->chunk_total + 3) / 4), 128, 0, e->stream>>>(w, b->d_qc); e->launches++; } else
#endif
    { AX_LAUNCH(e, k_qc, b->chunk_total, w, b->d_qc); }
    AX_LAUNCH(e, k_rows, b->frame_total, w);
    AX_LAUNCH(e, k_chunkout, b->chunk_total, w);
    if (streaming) { AX_LAUNCH(e, k_stream_commit, n, w); }
    AX_EVENT(b, 5);
    if (ax_launch_check(e)) return AXCTD_ERR_CUDA;
    b->ran = true;
    return AXCTD_OK;
}

extern "C" int axctd_batch_finish(axctd_batch* b) {
    if (!b || !b->ran) return AXCTD_ERR_STATE;
    if (b->finished) return AXCTD_OK;
    axctd_engine* e = b->eng;
    AxWave& w = b->w;
    const int n = b->n;
    AX_DEV(e);
    if (b->ran_nosync || b->ran_fused) {
        int32_t flags[8];
        if (ax_d2h(e, flags, w.flags, sizeof(flags)) || ax_d2h(e, b->h_st, w.st, sizeof(AxState) * n) || ax_sync(e)) return AXCTD_ERR_CUDA;
        const bool more = b->ran_nosync && (flags[AX_FLAG_MORE] || flags[AX_FLAG_DIRTY]);     // the fixed schedule was not enough for some drop
        const bool ovf = b->ran_fused && flags[AX_FLAG_FIXOVF];                              // more bits to re-evaluate than the list holds
        if (more || ovf) {
            b->force_sync = more || !b->ran_nosync; b->force_nofuse = ovf; b->n_fallbacks++;
            const int r = axctd_batch_run_async(b);
            b->force_sync = false; b->force_nofuse = false;
            if (r) return r;
            if (ax_d2h(e, b->h_st, w.st, sizeof(AxState) * n) || ax_sync(e)) return AXCTD_ERR_CUDA;
        }
    } else if (ax_d2h(e, b->h_st, w.st, sizeof(AxState) * n) || ax_sync(e)) return AXCTD_ERR_CUDA;
    memcpy(b->st.data(), b->h_st, sizeof(AxState) * n);
    for (int d = 0; d < n; ++d) {
        const AxDrop& dr = b->drops[d];
        const int64_t nf = b->st[d].status == 0 ? b->st[d].n_frames : 0;
        const int64_t nc = std::min<int64_t>(b->st[d].n_chunks, dr.chunk_cap);
        // (streaming: the rows of iterations closed by earlier runs are on the host already and do not change)
        const int64_t f0 = b->streaming ? std::min<int64_t>(b->rows_done[d], nf) : 0;
        if (nf > f0 && ax_d2h(e, b->h_row + dr.frame_base + f0, w.row + dr.frame_base + f0, sizeof(axctd_row) * (nf - f0))) return AXCTD_ERR_CUDA;
        if (b->streaming) b->rows_done[d] = nf;
        if (nc > 0 && ax_d2h(e, b->h_chunk + dr.chunk_base, w.chunk_out + dr.chunk_base, sizeof(axctd_chunk) * nc)) return AXCTD_ERR_CUDA;
    }
    if (ax_sync(e)) return AXCTD_ERR_CUDA;
#ifndef AXCTD_EMU
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev[0], b->ev[5]); b->ms_total = ms;
    cudaEventElapsedTime(&ms, b->ev[1], b->ev[2]); b->ms_filter = ms;
    cudaEventElapsedTime(&ms, b->ev[3], b->ev[4]); b->ms_tone = ms;
    for (int q = 0; q < 5; ++q) { cudaEventElapsedTime(&ms, b->ev[q], b->ev[q + 1]); b->ms_phase[q] = ms; }
#endif
    for (int d = 0; d < n; ++d) {
        const AxDrop& dr = b->drops[d];
        const AxCfg& c = e->cfgs[dr.cfg];
        AxState& st = b->st[d];
        axctd_drop_summary& sm = b->summary[d];
        if (st.n_uncertain > 0 && st.status == 0) { st.status = AXCTD_DROP_UNCERTAIN; }
        sm.status = st.status; sm.status_chunk = st.status_chunk;
        sm.numpoints = dr.n; sm.f_s = c.fs;
        sm.firstpulse400 = st.firstpulse400; sm.profstartind = st.profstartind;
        sm.firstpointtime = st.firstpointtime; sm.mean7500pwr = st.mean7500; sm.high_bit_scale = st.scale;
        sm.n_chunks = st.n_chunks; sm.first_demod_chunk = st.k0; sm.profile_chunk = st.k2;
        for (int q = 0; q < 3; ++q) { sm.header_read[q] = st.header_read[q]; sm.header_chunk[q] = st.header_chunk[q]; }
        sm.n_bits = st.nbits_total; sm.n_edges = st.nedges_total; sm.n_power = st.pcount;
        sm.n_frames = st.n_frames; sm.n_crossings = st.zc_count;
        sm.n_uncertain = st.n_uncertain; sm.n_chain_fixups = st.n_fixups;
        sm.n_guard_hits = st.n_unc_listed; sm.n_guard_confirmed = st.n_unc_resolved;
        sm.pcm_sum = st.sum; sm.pcm_ampl = st.ampl;
        sm.n_recheck = st.n_recheck; memcpy(&sm.win32_max_rel_err, &st.err32_bits, sizeof(float)); sm.n_frame_respec = st.n_frame_respec;
        memcpy(sm.frame_data, st.frame_data, sizeof(sm.frame_data));
        memcpy(sm.counter_found, st.counter_found, sizeof(sm.counter_found));
        sm.header_parsed[0] = st.header_parsed[0]; sm.header_parsed[1] = st.header_parsed[1];
        for (int q = 0; q < 4; ++q) {
            sm.zcoeff_used[q] = st.zc_used[q]; sm.tcoeff_used[q] = st.tc_used[q]; sm.ccoeff_used[q] = st.cc_used[q];
            sm.zcoeff[q] = st.md_z[q]; sm.tcoeff[q] = st.md_t[q]; sm.ccoeff[q] = st.md_c[q];
            sm.zcoeff_valid[q] = st.md_zv[q]; sm.tcoeff_valid[q] = st.md_tv[q]; sm.ccoeff_valid[q] = st.md_cv[q];
        }
        int64_t rows = 0, hex = 0;
        for (int k = 0; k < st.n_chunks && k < dr.chunk_cap; ++k) { rows += b->h_chunk[dr.chunk_base + k].n_rows; hex += b->h_chunk[dr.chunk_base + k].n_hex; }
        sm.n_rows = rows; sm.n_hex = hex;
    }
    b->finished = true;
    return AXCTD_OK;
}

extern "C" int axctd_batch_run(axctd_batch* b) {
    int r = axctd_batch_run_async(b);
    if (r) return r;
    return axctd_batch_finish(b);
}

// ============================================================ C ABI: streaming
// A growing recording decoded one run at a time (the shape of the reference's own loop: AXCTDprocessor.py:283-338 is
// written per 2 s iteration with a `keepgoing` flag).  Each drop of the batch was created with the most samples it can
// take; samples are appended as they arrive and every run decodes the iterations that have become complete
// (s + pointsperloop inside the data, so that :299-300 cannot cut them short), on top of the device state the previous
// runs left: per-sample work (tone block sums, filter, crossings, windows) covers the new samples only, per-bit work
// the new iterations only.  The normalisation of AXCTDprocessor.py:55-57 needs the whole file, so the caller fixes
// (dc, ampl) up front; the result equals the reference's for the recording normalised with those two numbers.
static void ax_stream_set_len(axctd_batch* b, int d, int64_t n) {
    AxDrop& dr = b->drops[d];
    dr.n = n; dr.n_raw = n;
    dr.nseg = (int32_t)((n + b->w.seg_len - 1) / b->w.seg_len);
    dr.nslab = (int32_t)((n + AX_STAT_SLAB - 1) / AX_STAT_SLAB);
    dr.ntb = (int32_t)(n / AX_TB);
}
extern "C" int axctd_batch_stream_begin(axctd_batch* b, const double* dc, const double* ampl) {
    if (!b || !dc || !ampl) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    if (b->streaming || b->ran) { e->err = "axctd_batch_stream_begin on a batch already in use"; return AXCTD_ERR_STATE; }
    for (int d = 0; d < b->n; ++d) {
        if (b->drops[d].xf_off >= 0) { e->err = "a recording that has to be halved cannot be streamed (sosfiltfilt runs backwards over the whole file)"; return AXCTD_ERR_ARG; }
        if (!(ampl[d] > 0.0) || !(dc[d] == dc[d])) { e->err = "bad normalisation"; return AXCTD_ERR_ARG; }
    }
    AX_DEV(e);
    std::vector<double> norm(2 * (size_t)b->n);
    for (int d = 0; d < b->n; ++d) { norm[2 * d] = dc[d]; norm[2 * d + 1] = ampl[d]; }
    if (ax_alloc_arr(b, &b->d_norm, 2 * (int64_t)b->n) || ax_h2d(e, b->d_norm, norm.data(), sizeof(double) * norm.size()) ||
        ax_zero(e, b->w.seg_cnt, sizeof(int32_t) * (size_t)b->w.nseg_total) || ax_zero(e, b->w.seg_unc, sizeof(int32_t) * (size_t)b->w.nseg_total) ||
        ax_sync(e)) return AXCTD_ERR_CUDA;
    b->cap_n.resize(b->n); b->rows_done.assign(b->n, 0);
    for (int d = 0; d < b->n; ++d) { b->cap_n[d] = b->drops[d].n_raw; ax_stream_set_len(b, d, 0); }
    b->streaming = true; b->stream_runs = 0; b->stream_closed = false;
    b->w.streaming = 1;
    return AXCTD_OK;
}
extern "C" int axctd_batch_stream_append(axctd_batch* b, int drop, const int16_t* pcm, int64_t n) {
    if (!b || !b->streaming || b->stream_closed || drop < 0 || drop >= b->n || n < 0 || (n > 0 && !pcm)) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    const int64_t have = b->drops[drop].n_raw;
    if (have + n > b->cap_n[drop]) { e->err = "recording longer than the batch was created for"; return AXCTD_ERR_CAPACITY; }
    AX_DEV(e);
    if (n > 0 && ax_h2d(e, b->d_pcm + b->drops[drop].pcm_off + have, pcm, sizeof(int16_t) * (size_t)n)) return AXCTD_ERR_CUDA;
    ax_stream_set_len(b, drop, have + n);
    return AXCTD_OK;
}
extern "C" int axctd_batch_stream_run(axctd_batch* b, int final_run) {
    if (!b || !b->streaming) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    if (b->stream_closed) { e->err = "the recording was closed by an earlier final run"; return AXCTD_ERR_STATE; }
    AX_DEV(e);
    b->w.streaming = final_run ? 2 : 1;
    if (ax_h2d(e, (void*)b->w.drop, b->drops.data(), sizeof(AxDrop) * (size_t)b->n)) return AXCTD_ERR_CUDA;
    b->in_stream_run = true;
    int r = axctd_batch_run_async(b);
    b->in_stream_run = false;
    if (!r) r = axctd_batch_finish(b);
    if (r) return r;
    b->stream_runs++;
    if (final_run) b->stream_closed = true;
    return AXCTD_OK;
}

extern "C" int axctd_batch_timing(axctd_batch* b, double* total_ms, double* filter_ms, double* tone_ms) {
    if (!b || !b->finished) return AXCTD_ERR_STATE;
    if (total_ms) *total_ms = b->ms_total;
    if (filter_ms) *filter_ms = b->ms_filter;
    if (tone_ms) *tone_ms = b->ms_tone;
    return AXCTD_OK;
}

extern "C" int axctd_batch_phase_ms(axctd_batch* b, double* ms5) {
    if (!b || !b->finished || !ms5) return AXCTD_ERR_STATE;
    for (int q = 0; q < 5; ++q) ms5[q] = b->ms_phase[q];
    return AXCTD_OK;
}

extern "C" int axctd_batch_summary(axctd_batch* b, int drop, axctd_drop_summary* out) {
    if (!b || !b->finished || drop < 0 || drop >= b->n || !out) return AXCTD_ERR_ARG;
    *out = b->summary[drop];
    return AXCTD_OK;
}

extern "C" int64_t axctd_batch_rows(axctd_batch* b, int drop, axctd_row* out, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    const int64_t nf = b->st[drop].status == 0 ? b->st[drop].n_frames : 0;
    if (!out) return nf;
    if (cap < nf) return -AXCTD_ERR_CAPACITY;
    if (nf) memcpy(out, b->h_row + b->drops[drop].frame_base, sizeof(axctd_row) * nf);
    return nf;
}

extern "C" int64_t axctd_batch_frames(axctd_batch* b, int drop, axctd_frame* out, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    const int64_t nf = b->st[drop].n_frames;
    if (!out) return nf;
    if (cap < nf) return -AXCTD_ERR_CAPACITY;
    if (nf && (ax_d2h(b->eng, out, b->w.frame + b->drops[drop].frame_base, sizeof(axctd_frame) * nf) || ax_sync(b->eng))) return -AXCTD_ERR_CUDA;
    return nf;
}

extern "C" int64_t axctd_batch_chunks(axctd_batch* b, int drop, axctd_chunk* out, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    const AxDrop& dr = b->drops[drop];
    const int64_t nc = std::min<int64_t>(b->st[drop].n_chunks, dr.chunk_cap);
    if (!out) return nc;
    if (cap < nc) return -AXCTD_ERR_CAPACITY;
    if (nc) memcpy(out, b->h_chunk + dr.chunk_base, sizeof(axctd_chunk) * nc);
    return nc;
}

extern "C" int64_t axctd_batch_bits(axctd_batch* b, int drop, uint8_t* bits, double* conf, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    const int64_t nb = b->st[drop].nbits_total;
    if (!bits && !conf) return nb;
    if (cap < nb) return -AXCTD_ERR_CAPACITY;
    const int64_t base = b->drops[drop].edge_base;
    // bits of chunk k sit at bit_off[k]: contiguous over the drop
    if (bits && nb && ax_d2h(b->eng, bits, b->w.bit + base, (size_t)nb)) return -AXCTD_ERR_CUDA;
    if (conf && nb) {
#ifndef AXCTD_EMU
        if (!b->mags_full) {      // the run kept magnitudes only where it needed them: produce the rest now (two-step form of the later iterations)
            k_emit_chunk<<<dim3((unsigned)b->chunk_cap_max, (unsigned)b->n), 128, 0, b->eng->stream>>>(b->w, 3);
            k_bits_chunk<<<(unsigned)b->chunk_total, 128, 0, b->eng->stream>>>(b->w, 1, 2);
            if (ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
            b->mags_full = true;
        }
#endif
        // demodulate.py:102,110: conf = |S2| * high_bit_scale / |S1| with the scale in force when the bit was demodulated
        // (ax_bits_decide's arithmetic, on the magnitudes the decision was made from)
        std::vector<double> p1((size_t)nb), p2((size_t)nb);
        if (ax_d2h(b->eng, p1.data(), b->w.a1 + base, sizeof(double) * nb) || ax_d2h(b->eng, p2.data(), b->w.a2 + base, sizeof(double) * nb) ||
            ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
        const AxState& st = b->st[drop];
        const double s0 = b->eng->cfgs[b->drops[drop].cfg].scale0;
        for (int64_t j = 0; j < nb; ++j) conf[j] = ax_div(ax_mul(p2[j], j >= st.scale_switch_bit ? st.scale : s0), p1[j]);
    }
    if (ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
    return nb;
}

extern "C" int64_t axctd_batch_edges(axctd_batch* b, int drop, int64_t* edges, double* r400, double* r7500, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    const int64_t ne = b->st[drop].nedges_total;
    if (!edges && !r400 && !r7500) return ne;
    if (cap < ne) return -AXCTD_ERR_CAPACITY;
    const int64_t base = b->drops[drop].edge_base;
    if (edges && ne) {
        std::vector<int32_t> tmp(ne);
        if (ax_d2h(b->eng, tmp.data(), b->w.edge_idx + base, sizeof(int32_t) * ne) || ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
        for (int64_t i = 0; i < ne; ++i) edges[i] = tmp[i];
    }
    if ((r400 || r7500) && ne) {
        // edges hold the index of the power sample whose levels they take: gather on the host
        const int64_t np = b->drops[drop].pw_cap, pb = b->drops[drop].pw_base;
        std::vector<int32_t> sl(ne);
        std::vector<double> l4(np > 0 ? np : 1), l7(np > 0 ? np : 1);
        if (ax_d2h(b->eng, sl.data(), b->w.lvl_slot + base, sizeof(int32_t) * ne) ||
            (np > 0 && ax_d2h(b->eng, l4.data(), b->w.r400 + pb, sizeof(double) * np)) ||
            (np > 0 && ax_d2h(b->eng, l7.data(), b->w.r7500m + pb, sizeof(double) * np)) || ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
        const double nan = std::numeric_limits<double>::quiet_NaN();
        for (int64_t i = 0; i < ne; ++i) {
            const bool ok = sl[i] >= 0 && sl[i] < np;
            if (r400) r400[i] = ok ? l4[sl[i]] : nan;
            if (r7500) r7500[i] = ok ? l7[sl[i]] : nan;
        }
    }
    return ne;
}

extern "C" int64_t axctd_batch_power(axctd_batch* b, int drop, int64_t* power_inds, double* r400, double* r7500, int64_t cap) {
    if (!b || !b->finished || drop < 0 || drop >= b->n) return -AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    const int64_t np = b->st[drop].pcount;
    if (!power_inds && !r400 && !r7500) return np;
    if (cap < np) return -AXCTD_ERR_CAPACITY;
    const int64_t base = b->drops[drop].pw_base;
    if (power_inds && np && ax_d2h(b->eng, power_inds, b->w.pw_ind + base, sizeof(int64_t) * np)) return -AXCTD_ERR_CUDA;
    if (r400 && np && ax_d2h(b->eng, r400, b->w.r400 + base, sizeof(double) * np)) return -AXCTD_ERR_CUDA;
    if (r7500 && np && ax_d2h(b->eng, r7500, b->w.r7500 + base, sizeof(double) * np)) return -AXCTD_ERR_CUDA;
    if (ax_sync(b->eng)) return -AXCTD_ERR_CUDA;
    return np;
}

// ============================================================ bench / test tooling
extern "C" int axctd_synth_fill(axctd_batch* b, int drop, const axctd_synth_desc* ds) {
    if (!b || !ds || drop < 0 || drop >= b->n || ds->n_total != b->drops[drop].n_raw || !ds->bits || !ds->gate || !ds->parity) return AXCTD_ERR_ARG;
    axctd_engine* e = b->eng;
    AX_DEV(e);
    AxSynth g;
    g.n_total = ds->n_total; g.n0 = ds->n0; g.tone_start = ds->tone_start; g.fs = ds->fs;
    g.key1 = ds->key1; g.key2 = ds->key2; g.nscale = ds->nscale; g.gain = ds->gain; g.tone_amp = ds->tone_amp;
    for (int q = 0; q < 9; ++q) g.sin_coef[q] = ds->sin_coef[q];
    g.nslots = ds->nslots;
    uint8_t* dev = nullptr;
    if (ax_alloc_arr(b, &dev, 3 * ds->nslots + 16)) return AXCTD_ERR_CUDA;
    if (ax_h2d(e, dev, ds->bits, (size_t)ds->nslots) || ax_h2d(e, dev + ds->nslots, ds->gate, (size_t)ds->nslots) ||
        ax_h2d(e, dev + 2 * ds->nslots, ds->parity, (size_t)ds->nslots)) return AXCTD_ERR_CUDA;
    g.bits = dev; g.gate = dev + ds->nslots; g.par = dev + 2 * ds->nslots;
    g.out = b->d_pcm + b->drops[drop].pcm_off;
    int64_t launches_before = e->launches;
    AX_LAUNCH(e, k_synth, ds->n_total, g);
    e->launches = launches_before;          // tooling, not part of the decode path
    if (ax_sync(e)) return AXCTD_ERR_CUDA;
    b->ran = false;
    return AXCTD_OK;
}

extern "C" int axctd_calib_eval(axctd_engine* e, const double* cond, const double* temp, const double* pres, int n,
                                const double* coeff4, double* sp, double* poly) {
    if (!e || !cond || !temp || !pres || !sp || n <= 0 || (coeff4 && !poly)) return AXCTD_ERR_ARG;
    AX_DEV(e);
    std::vector<double> in(3 * (size_t)n + 4, 0.0), out(2 * (size_t)n);
    memcpy(in.data(), cond, sizeof(double) * n); memcpy(in.data() + n, temp, sizeof(double) * n);
    memcpy(in.data() + 2 * (size_t)n, pres, sizeof(double) * n);
    if (coeff4) memcpy(in.data() + 3 * (size_t)n, coeff4, sizeof(double) * 4);
    void *d_in = nullptr, *d_out = nullptr;
    int bad = ax_alloc(e, &d_in, in.size() * sizeof(double)) || ax_alloc(e, &d_out, out.size() * sizeof(double));
    const int64_t launches_before = e->launches;
    if (!bad) {
        bad = ax_h2d(e, d_in, in.data(), in.size() * sizeof(double));
        if (!bad) { AX_LAUNCH(e, k_calib_eval, (int64_t)n, (const double*)d_in, (double*)d_out, coeff4 ? 1 : 0); }
        bad = bad || ax_d2h(e, out.data(), d_out, out.size() * sizeof(double)) || ax_sync(e);
    }
    e->launches = launches_before;          // tooling, not part of the decode path
    ax_free(e, d_in); ax_free(e, d_out);
    if (bad) return AXCTD_ERR_CUDA;
    memcpy(sp, out.data(), sizeof(double) * n);
    if (poly) memcpy(poly, out.data() + n, sizeof(double) * n);
    return AXCTD_OK;
}

extern "C" int axctd_batch_download(axctd_batch* b, int drop, int16_t* pcm, int64_t n) {
    if (!b || drop < 0 || drop >= b->n || !pcm || n != b->drops[drop].n_raw) return AXCTD_ERR_ARG;
    AX_DEV(b->eng);
    if (ax_d2h(b->eng, pcm, b->d_pcm + b->drops[drop].pcm_off, sizeof(int16_t) * n) || ax_sync(b->eng)) return AXCTD_ERR_CUDA;
    return AXCTD_OK;
}
