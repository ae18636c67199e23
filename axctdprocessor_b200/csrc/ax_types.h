// ax_types.h -- device-side data model of the AXCTD engine.
//
// All per-batch state lives in flat device arrays described by AxWave; every
// kernel receives the AxWave by value.  The same header compiles for the GPU
// (nvcc) and, with -DAXCTD_EMU, for the test-only host emulation used to debug
// the algorithm without a GPU (tests/emu; never loaded by the product).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/axctd.h"

#if defined(__CUDACC__) && !defined(AXCTD_EMU)
#define AX_HD __host__ __device__ __forceinline__
#define AX_HDN __host__ __device__
#else
#define AX_HD inline
#define AX_HDN
#endif

// ---- IEEE operations that must never be contracted into FMAs -------------
// (exact restatement of scipy's C sosfilt, which is compiled without FMA)
AX_HD double ax_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;   // host objects are built with -ffp-contract=off
#endif
}
AX_HD double ax_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
AX_HD double ax_sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, -b);
#else
    return a - b;
#endif
}
AX_HD double ax_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
AX_HD double ax_fma(double a, double b, double c) { return fma(a, b, c); }
AX_HD double ax_nan() { return nan(""); }

#define AX_TILE 64          // crossings per walk tile
#define AX_ZQ 1024          // samples per entry of the coarse crossing index
#define AX_ZQ_SHIFT 10
#define AX_MAXSEC AXCTD_MAX_SECTIONS
#define AX_PEND 8           // pending bit-windows per thread in the filter pass
#define AX_STAT_SLAB 16384  // samples per stats work item
#define AX_TB 256           // samples per tone block (ax_toneblock_item)
#define AX_TONE_ROT 40      // most tone blocks a window can span (0.1 s at 96 kHz is 37.5 blocks)
#define AX_POW10_LEN 1099
#define AX_UNC_CAP 256      // guard-band samples listed per drop (ax_unc_push); more are only counted

// fp32 phasor table of the bit windows (ax_window32): cos, sin of theta_mark * k and theta_space * k
#define AX_WIN_TAPS 48
struct alignas(16) AxF4 { float x, y, z, w; };
struct AxWinTab { AxF4 t[AX_WIN_TAPS]; };

// Per rate-class constants (reference AXCTDprocessor.py:117-182, 212-262).
struct AxCfg {
    double fs;
    int64_t fs2;                 // round(2*fs): the walk compares |d*2*bitrate - fs2| (demodulate.py:92)
    int32_t n_power, d_pcm, npcm, chunk_len, pad, inset, bitrate, nsec;
    double sos[AX_MAXSEC][6];
    int32_t warm;                // warm-up overlap of the continuous filter pass (samples)
    int32_t head;                // exact zero-state prefix recomputed per chunk (samples)
    int32_t head_zc_cap;         // crossings kept per chunk head
    double min_r400, min_dr7500, min_r400_inprof, min_dr7500_inprof;
    double trig_from, trig_to, scale0;
    double zc[4], tc[4], cc[4], tlims[2], slims[2];
    int64_t off_4p5, off_5p5, off_trig_from, off_trig_to;          // int(f_s*x)
    int64_t h1s, h1e, h2s, h2e, h3s, h3e, half;                    // AXCTDprocessor.py:447-456
    const double* tone_cs;       // [n_power][6]
    const double* tone_tab8;     // [AX_TB][8] phasors of a tone block, columns 6 and 7 zero (B operand of k_stats_tones_mma)
    const uint32_t* tone_tabi;   // the same phasors as signed base-256 digits of round(p * 2^45), in the B-fragment layout of k_stats_tones_imma
    const double* tone_soa;      // [6][n_power] the same table, one array per component (coalesced reads in ax_tonewin_partial)
    const double* lut;           // [lut_len]
    const double* hist_edges;    // [n_hist_edges]
    const double* hist_centers;  // [n_hist_edges-1]
    const double* gtab;          // [g_len][4] mark/space window responses: |S_f(i)| = |sum_d u[i+npcm-d] * G_f[d]| (ax_gwin_*)
    const double* gcum;          // [g_len][4] running sums of gtab (response to the constant -dc/ampl term)
    const double* pow10;         // [AX_POW10_LEN] python's 10**ex as a float for ex = -99 .. 999 (parse.py:278), entry ex + 99
    int32_t g_len, pad_g;
    int32_t lut_len, n_hist_edges;
    double tone_tsum[6];         // sum over the whole window of tone_cs (response to the constant -dc/ampl term)
    double tone_rot[AX_TONE_ROT][6];   // tone_cs rows AX_TB * j: phasors that carry a tone block sum j blocks into a window
    // /2 decimation (AXCTDprocessor.py:60-62): forward-backward Chebyshev SOS with odd padding
    int32_t decimate, dnsec, dpad, dwarm;
    double dsos[AX_MAXSEC][6], dzi[AX_MAXSEC][2];
    AxWinTab win_tab;
};

struct AxDrop {
    int64_t pcm_off, n;          // sample offset into the batch PCM buffer, sample count (after any decimation)
    int64_t n_raw;               // samples of the recording as uploaded
    int64_t xf_off, fwd_off;     // decimating drops: offsets of the halved signal / forward-pass scratch (doubles), else -1
    int32_t dseg_base, ndseg;    // segments of the two decimation passes
    int32_t cfg;
    int32_t seg_base, nseg;      // segments of the continuous filter pass
    int32_t slab_base, nslab;    // stats work items
    int64_t tb_base; int32_t ntb, pad_tb;   // tone blocks: floor(n / AX_TB) full blocks
    int64_t zc_base, zc_cap;     // dense crossing arrays
    int64_t zq_base;             // coarse index of the dense crossings (AxWave::zc_q), n / AX_ZQ + 2 entries
    int32_t tile_base, tile_cap;
    int32_t chunk_base, chunk_cap;
    int64_t edge_base, edge_cap; // bit / edge arrays
    int32_t pw_base, pw_cap;     // power samples
    int32_t frame_base, frame_cap;
    int32_t pad0;
};

// One run() iteration (reference AXCTDprocessor.py:283-338).
struct AxChunk {
    int64_t s, e;                // demodbufferstartind, e
    int64_t spec_last;           // last bit edge (global PCM index) predicted from the continuous pass
    int64_t true_last;           // last bit edge with the exact zero-state head
    int64_t g_first;             // dense ordinal of the first bit edge taken from the continuous pass, -1 if none
    int64_t q_last;              // dense ordinal of the chunk's last usable crossing
    int64_t bit_off, edge_off;   // offsets into the drop's bit / edge arrays
    int64_t first_edge;          // first bit edge (global PCM index)
    int64_t merge_pos;           // dense ordinal where the chunk's walk joins the canonical walk, -1 if it never does
    int32_t n_pre, pad_pre;      // edges of the continuous part before merge_pos (stepped explicitly)
    int64_t scan_from, scan_end; // frame scan of this iteration: bit position it starts from / leaves to the next one
    int32_t scan_cnt, pad_scan;  // frames it finds
    int32_t pw_off, np;          // power samples of this iteration
    int32_t n_head_edges;        // bit edges found inside the exact head
    int32_t n_edges;             // total bit edges (bits = n_edges - 1), 0 if not demodulated
    int32_t status;              // self.status after the iteration
    int32_t err;                 // AXCTD_DROP_* raised by this chunk
    int32_t n_rows, n_hex;
    int32_t frame_begin, frame_end;   // frames parsed in this iteration
    double scale;
    double mean7500;             // mean7500pwr in force when the chunk was demodulated (NaN before)
    int64_t profstart;           // self.profstartind after this iteration's detectors (the timeout branch of
                                 // AXCTDprocessor.py:404-408 can move it after status 2 was reached)
};

// Sequential per-drop state (the attributes of reference class AXCTD_Processor).
struct AxState {
    int32_t status, status_chunk;
    int64_t sum;
    int32_t vmax, vmin;          // max / min sample (k_stats_tones); vmin stays INT_MAX on the generic path
    int32_t ampl;
    int32_t n_uncertain;         // guard-band samples whose sign could not be confirmed (ax_unc_resolve_item / ax_unc_fin_item)
    int32_t n_unc_listed;        // guard-band samples seen by the filter passes (the first AX_UNC_CAP are listed)
    int32_t n_unc_relevant;      // listed ones that lie where a demodulated iteration takes crossings
    int32_t n_unc_resolved;      //   ... and whose sign the exact zero-state recomputation confirmed
    int32_t n_recheck;           // bit windows re-evaluated in double precision (ax_gwin_*)
    int32_t err32_bits;          // max relative |a32 - a64| / a64 seen at re-evaluated windows (float bits)
    double dc, inv_ampl, ampl_d;
    double dc_raw, ampl_raw;     // decimating drops: statistics of the raw recording (dc / ampl above become 0 / 1)
    int64_t zc_count;
    // state machine
    int32_t sm_status;           // self.status
    int32_t n_chunks;            // run() iterations
    int32_t n_fixed;             // iterations of the fixed (status 0) grid that were planned
    int32_t k0, k2, km;          // first demod chunk, profile chunk, chunk where mean7500pwr was set
    int32_t pcount;              // len(power_inds)
    int32_t next_sm_chunk;       // next iteration the level state machine will process
    int64_t firstpulse400, profstartind;
    double firstpointtime, mean7500;
    // chain
    int32_t chain_from;          // first chunk whose start is not yet final
    int32_t chain_dirty;         // set by verify when a mis-speculation was repaired
    int32_t n_fixups;
    int32_t chain_end;           // chain finished (file end or error)
    // demod bookkeeping
    double scale;
    int32_t k1;                  // iteration at the end of which the scale changed
    int32_t par_levels;          // 1: smoothing may be evaluated per power sample in parallel
    int64_t scale_switch_bit;    // first bit decided with the calibrated scale
    int32_t searching;           // still looking for the 400 Hz pulse on the fixed grid
    int32_t pad2;
    int32_t header_read[3], header_chunk[3];
    int64_t nbits_total, nedges_total;
    int64_t n_frames, n_rows, n_hex;
    int32_t n_frame_respec, pad4; // iterations whose speculative frame scan had to be redone
    int32_t header_parsed[2];
    uint16_t frame_data[2][72];
    uint8_t counter_found[2][72];
    double zc_used[4], tc_used[4], cc_used[4];
    // merged header metadata (AXCTDprocessor.py:505-524): metadata['?coeff'] and metadata['?coeff_valid']
    double md_z[4], md_t[4], md_c[4];
    uint8_t md_zv[4], md_tv[4], md_cv[4];
    int32_t pad5;
    // streaming decode (axctd_batch_stream_*): what the previous run left final
    int32_t seg_done;            // leading segments of the continuous pass whose crossing records are final
    int32_t k_done;              // leading run() iterations whose edges / bits are final
    int64_t tb_done;             // leading tone blocks whose sums are final
};

struct AxWave {
    int32_t n_drops, n_cfg;
    const AxCfg* cfg;
    const AxDrop* drop;
    AxState* st;
    const int16_t* pcm;
    double* xf; double* fwd;     // decimated recordings (already normalised) and the forward-pass scratch
    int32_t dseg_len, only_xf;   // decimation segment length; only_xf: generic sample-stream kernels skip int16 drops
    // continuous filter pass
    int32_t seg_len, seg_cap, nseg_total, nslab_total;
    const int32_t* seg_drop;     // segment -> drop
    const int32_t* slab_drop;    // stats slab -> drop
    int32_t* seg_cnt; int64_t* seg_off; int64_t* blk_sum;  // per segment / per block of 128 segments
    int32_t* seg_unc; int32_t* head_unc;                    // guard-band samples per segment / per chunk head
    int64_t* unc_list;                                      // [n_drops][AX_UNC_CAP][2]: sample | neg << 32 | (chunk + 1) << 33, chunk start
    int32_t* rec_idx; float* rec_a1; float* rec_a2;         // [nseg_total * seg_cap]
    int32_t* zc_idx; float* zc_a1; float* zc_a2;            // dense, per drop at zc_base
    uint8_t* zc_nx;                                         // per crossing: walk step
    int32_t* zc_q;                                          // per AX_ZQ samples: crossings with index below AX_ZQ * j (k_compact_warp -> k_chain_warp)
    uint64_t* tile_mask; uint32_t* tile_map;                // [tile][4] visited masks, [tile] exit offsets (ax_tiles_item)
    uint64_t* cmask; int32_t* crank;                        // [tile] canonical walk: visited mask, visited count before the tile
    // chunks
    AxChunk* chunk;
    int32_t* head_idx; float* head_a1; float* head_a2;      // [chunk][head_zc_cap_max] crossings of the zero-state heads
    int32_t* head_cnt;                                      // [chunk] how many (-1: did not fit)
    int32_t head_zc_cap_max, pad_head;
    // tone powers
    double* pw_raw;              // [3][pw_total]
    double* pw_sm;               // [3][pw_total]
    double* r400; double* r7500; // [pw_total]
    int64_t* pw_ind;             // [pw_total] power_inds
    int32_t* tone_rng;           // [n_drops][2] power-sample range the current tone launch evaluates (k_tone_range)
    double* tone_acc;            // [pw_total][6] window sums before normalisation (k_tone_windows -> k_tone_mag)
    int32_t pw_total;
    double* tb_sum;              // [tb_total][6] raw-sample tone sums of every aligned block of AX_TB samples
    int32_t ntb_max, pad1;
    // bits / edges
    int32_t* edge_idx; int32_t* lvl_slot;                   // per edge: PCM index; power sample whose levels it takes (-1: none)
    double* r7500m;                                         // [pw_total] r7500 - mean7500pwr of the iteration the sample belongs to
    uint8_t* bit; double* a1; double* a2;                   // per bit
    uint32_t* bitw; uint32_t* validw;                       // packed bits / frame-candidate mask (32 per word)
    // frames
    axctd_frame* frame;
    axctd_row* row;              // [frame_total] compact results (ax_row_item)
    axctd_chunk* chunk_out;      // [chunk_total] public per-iteration records (ax_chunkout_item)
    double guard;
    double bit_tol;              // |p1 - p2| <= bit_tol * max(p1, p2): the bit is re-decided from a double-precision window
    double hist_tol;             // conf within hist_tol (relative) of a histogram bin edge: re-evaluated before the scale calibration
    int32_t bitfix_all;          // test hook: re-evaluate every bit window in double precision
    int32_t probe;               // timing probe of k_demod_fused (builds with -DAX_DEMOD_PROBE only; option "demod_probe", results are wrong): 1 = window sums skipped, 2 = crossings not processed at all
    int32_t tone_direct;
    int32_t tone_complement;     // k_tone_windows_mma: ragged ends above half a block as the block minus its complement
    int32_t pa_lo, pa_hi;        // fixed-grid chunk range of the current detection round
    int32_t force_exact;
    int32_t streaming;           // 0: whole recordings; 1: a growing recording, only iterations that are complete are
                                 // decoded (AXCTDprocessor.py:293-304 with the end of the file still unknown); 2: its last run
    int32_t* flags;              // [0] = any chain dirty, [1] = any capacity error, [2] = another search round, [3] = fix list full, [4] = fix list length
    // bit decisions made while the edges are emitted (k_emit_chunk, fused modes): iterations before k0 + nk_full keep
    // the two-step form (they feed the scale calibration); later ones write bits directly and list the few bits whose
    // decision needs a double-precision window (k_bits_recheck)
    int32_t nk_full, pad_nk;
    int64_t* fix_list; int64_t fix_cap;
};

// cos, sin of theta400, theta7500, thetadead for the AX_TB taps of a tone block (kernel parameter of k_stats_tones)
struct AxToneTab { double t[AX_TB][6]; };
// the samples of one drop: int16 as uploaded, or the halved double-precision signal
struct AxSrc { const int16_t* x; const double* xf; };
AX_HD double ax_get(const AxSrc& s, int64_t n) { return s.xf ? s.xf[n] : (double)s.x[n]; }

AX_HD bool ax_tone_blocked_ok(const AxCfg& c) { return c.n_power >= 2 * AX_TB && c.n_power / AX_TB < AX_TONE_ROT; }

AX_HD AxSrc ax_src(const AxWave& w, const AxDrop& dr) {
    AxSrc s;
    s.x = w.pcm + dr.pcm_off;
    s.xf = dr.xf_off >= 0 ? w.xf + dr.xf_off : nullptr;
    return s;
}

// the non-zero fields of a fresh state record (k_init clears the rest)
AX_HD void ax_state_defaults(AxState& st, const AxCfg& c) {
    st.ampl = -2147483647 - 1; st.vmax = -2147483647 - 1; st.vmin = 0x7fffffff;
    st.k0 = st.k2 = st.km = st.k1 = -1;
    st.firstpulse400 = -1; st.profstartind = -1; st.firstpointtime = -1.0; st.mean7500 = ax_nan();
    st.status_chunk = -1;
    for (int q = 0; q < 3; ++q) st.header_chunk[q] = -1;
    st.scale = c.scale0;
    for (int q = 0; q < 4; ++q) { st.zc_used[q] = c.zc[q]; st.tc_used[q] = c.tc[q]; st.cc_used[q] = c.cc[q]; }
}

#define AX_FLAG_DIRTY 0
#define AX_FLAG_CAP 1
#define AX_FLAG_MORE 2
#define AX_FLAG_FIXOVF 3
#define AX_FLAG_FIXCNT 4

AX_HD void ax_raise(AxState& st, int code, int chunk) {
    if (st.status == 0) { st.status = code; st.status_chunk = chunk; }
}
