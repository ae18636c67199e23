// ax_proto.h -- protocol layer: header decode, frame sync + CRC, calibration, QC.
//
//   ax_header_item   header windows, trim_header, parse_header   AXCTDprocessor.py:472-501, parse.py:157-183, 197-245
//   ax_frames_item   profile trim + greedy frame sync + CRC       AXCTDprocessor.py:540-557, 617-621, parse.py:41-92, 310-322
//   ax_calib_item    LUT, C*60/4096, cubics, PSS-78, rounding, QC AXCTDprocessor.py:559-574, parse.py:103-134, 297-301
//   ax_qc_item       per-iteration spike filter                   AXCTDprocessor.py:576-613
#pragma once
#include "ax_levels.h"

// parse.py:310-322: remainder of the 32-bit word modulo x^6+x^5+x^2+1 is zero
AX_HD bool ax_crc_ok(uint32_t wd) {
#pragma unroll
    for (int k = 0; k < 26; ++k) if (wd & (0x80000000u >> k)) wd ^= (0x65u << (25 - k));
    return wd == 0;
}

AX_HD uint32_t ax_word32(const uint8_t* b) {
    uint32_t wd = 0;
    for (int i = 0; i < 32; ++i) wd = (wd << 1) | (b[i] ? 1u : 0u);
    return wd;
}

// One header slot (0 = second transmission, 1 = third) of one drop, in three steps so that the CUDA build can stage
// the bit window in shared memory (k_headers_warp): reset, the window of one candidate iteration, trim + parse.
AX_HD void ax_header_reset(AxState& st, int slot, int first, int step) {
    if (first == 0) { st.header_read[1 + slot] = 0; st.header_chunk[1 + slot] = -1; st.header_parsed[slot] = 0; }
    for (int q = first; q < 72; q += step) { st.frame_data[slot][q] = 0; st.counter_found[slot][q] = 0; }
}
// iteration k: 0 = its buffer does not span the header yet, 1 = bits [a, a + n) of the drop are the window
// (AXCTDprocessor.py:472-478), < 0 = -AXCTD_DROP_* (the reference's IndexError)
AX_HD int ax_header_window(const AxWave& w, const AxDrop& dr, const AxCfg& c, const AxState& st, int slot, int k, int64_t* a_out, int64_t* n_out) {
    const AxChunk* ch = w.chunk + dr.chunk_base;
    const int32_t* I = w.edge_idx + dr.edge_base;
    const int64_t ps = st.firstpulse400 + (slot ? c.h3s : c.h2s), pe = st.firstpulse400 + (slot ? c.h3e : c.h2e);
    if (ch[k].n_edges <= 0) return 0;
    const int64_t ni = ch[k].edge_off + ch[k].n_edges, nb = ch[k].bit_off + ch[k].n_edges - 1;
    if (!((int64_t)I[0] <= ps && (int64_t)I[ni - 1] >= pe)) return 0;
    const int64_t a = ax_first_ge(I, ni, ps - c.half), b = ax_last_le(I, ni, pe + c.half);
    if (a < 0 || b < 0) return -AXCTD_DROP_TRIM_INDEX;
    const int64_t hi = b < nb ? b : nb;
    *a_out = a; *n_out = hi > a ? hi - a : 0;
    return 1;
}
// trim_header (parse.py:157-183) and parse_header (parse.py:197-245) on the window's bits; false: fewer than 72 frames
// are left after the trim, the next iteration tries again (AXCTDprocessor.py:481)
AX_HD bool ax_header_parse(AxState& st, int slot, int k, const uint8_t* bits, int64_t n) {
    int64_t last_pulse = 0; int ones25 = 0; int run = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int bv = (i < 25) ? 1 : bits[i];
        if (bv) { ++ones25; ++run; if (i > 10 && run >= 8) last_pulse = i; } else run = 0;
        if (i > 24) {
            const int back = (i - 25 < 25) ? 1 : bits[i - 25];
            if (back) --ones25;
            if (i >= 400 && ones25 <= 20) break;
        }
    }
    int64_t hn = n - last_pulse;
    if (hn > 32 * 75) hn = 32 * 75;
    if (hn < 72 * 32) return false;
    int lastframe = -1; int64_t s = 0;
    uint8_t fb[32];
    while (lastframe < 71 && s < hn - 32) {
        for (int q = 0; q < 32; ++q) { const int64_t p = last_pulse + s + q; fb[q] = (p < 25) ? 1 : bits[p]; }
        if (!(fb[0] == 1 && fb[1] == 0) || !ax_crc_ok(ax_word32(fb))) { ++s; continue; }
        int cur = 0;
        if (fb[2] & fb[3] & fb[4] & fb[5] & fb[6]) { cur = (fb[7] << 2 | fb[8] << 1 | fb[9]) + 64; }
        else { for (int q = 2; q < 10; ++q) cur = (cur << 1) | fb[q]; }
        if (cur <= 71) {
            st.counter_found[slot][cur] = 1; lastframe = cur;
            uint16_t v = 0; for (int q = 10; q < 26; ++q) v = (uint16_t)((v << 1) | fb[q]);
            st.frame_data[slot][cur] = v;
        }
        s += 32;
    }
    st.header_parsed[slot] = 1; st.header_read[1 + slot] = 1; st.header_chunk[1 + slot] = k;
    return true;
}
AX_HDN inline void ax_header_item(const AxWave& w, int64_t item) {
    const int d = (int)(item >> 1), slot = (int)(item & 1);
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    ax_header_reset(st, slot, 0, 1);
    if (st.sm_status < 1 || st.nedges_total == 0) return;
    const uint8_t* B = w.bit + dr.edge_base;
    const int klast = (st.k2 >= 0) ? st.k2 : st.n_chunks - 1;
    for (int k = st.k0; k <= klast && k < st.n_chunks; ++k) {
        int64_t a = 0, n = 0;
        const int r = ax_header_window(w, dr, c, st, slot, k, &a, &n);
        if (r < 0) { ax_raise(st, -r, k); return; }
        if (r == 0) continue;
        if (ax_header_parse(st, slot, k, B + a, n)) return;
    }
}

// ---- header text -> calibration coefficients and their merge (parse.py:272-283, AXCTDprocessor.py:505-535) ----
// A coefficient is the 12 hex characters of three consecutive frames read as text with B -> '+' and D -> '-':
// int(chars[0:9]) / 1E7 * 10**int(chars[9:12]) in Python arithmetic.  int() takes an optional sign and at least one
// decimal digit (any other character -- A, C, E, F or a misplaced sign -- is the ValueError of parse.py:278); the
// division is one IEEE division; 10**ex is an exact integer turned into a float for ex >= 0 and libm's pow(10., ex)
// for ex < 0, both tabulated by the host (AxCfg::pow10).
AX_HD bool ax_py_int_nibbles(const uint8_t* nib, int len, int64_t* out) {
    int i = 0, sign = 1;
    if (nib[0] == 0xB || nib[0] == 0xD) { if (nib[0] == 0xD) sign = -1; i = 1; }
    if (i >= len) return false;
    int64_t v = 0;
    for (; i < len; ++i) { if (nib[i] > 9) return false; v = v * 10 + nib[i]; }
    *out = sign * v;
    return true;
}
AX_HD bool ax_coeff_from_frames(const uint16_t* f3, const double* pow10, double* out) {
    uint8_t nib[12];
    for (int q = 0; q < 12; ++q) nib[q] = (uint8_t)((f3[q >> 2] >> (12 - 4 * (q & 3))) & 0xF);
    int64_t mant, ex;
    if (!ax_py_int_nibbles(nib, 9, &mant) || !ax_py_int_nibbles(nib + 9, 3, &ex)) return false;
    *out = ax_mul(ax_div((double)mant, 1e7), pow10[ex + 99]);
    return true;
}
AX_HDN inline void ax_merge_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    const double zd[4] = {1, 1, 1, 1}, td[4] = {0, 1, 0, 0};       // parse.py:187-192 defaults
    for (int q = 0; q < 4; ++q) {
        st.md_z[q] = zd[q]; st.md_t[q] = td[q]; st.md_c[q] = td[q];
        st.md_zv[q] = st.md_tv[q] = st.md_cv[q] = 0;
        st.zc_used[q] = c.zc[q]; st.tc_used[q] = c.tc[q]; st.cc_used[q] = c.cc[q];
    }
    bool any = false;
    for (int slot = 0; slot < 2; ++slot) {
        if (!st.header_parsed[slot]) continue;
        any = true;
        const uint8_t* cf = st.counter_found[slot];
        const uint16_t* fd = st.frame_data[slot];
        for (int set = 0; set < 3; ++set) {                         // parse.py:272: t, c, z
            const int top = set == 0 ? 33 : set == 1 ? 45 : 21;    // parse.py:258-270
            double* co = set == 0 ? st.md_t : set == 1 ? st.md_c : st.md_z;
            uint8_t* va = set == 0 ? st.md_tv : set == 1 ? st.md_cv : st.md_zv;
            for (int i = 0; i < 4; ++i) {
                const int f0 = top - 3 * i;
                if (cf[f0] && cf[f0 + 1] && cf[f0 + 2]) {
                    double v;
                    if (!ax_coeff_from_frames(fd + f0, c.pow10, &v)) { ax_raise(st, AXCTD_DROP_HEADER_VALUE, st.header_chunk[1 + slot]); return; }
                    co[i] = v; va[i] = 1;
                }
            }
        }
    }
    if (any) {                                                      // AXCTDprocessor.py:529-535
        const int tv = st.md_tv[0] + st.md_tv[1] + st.md_tv[2] + st.md_tv[3];
        const int cv = st.md_cv[0] + st.md_cv[1] + st.md_cv[2] + st.md_cv[3];
        if (tv == 4) for (int q = 0; q < 4; ++q) st.tc_used[q] = st.md_t[q];
        if (cv == 4) for (int q = 0; q < 4; ++q) st.cc_used[q] = st.md_c[q];
        if (tv == 4) for (int q = 0; q < 4; ++q) st.zc_used[q] = st.md_z[q];      // (sic) guarded by the T flag
    }
}

// 32 demodulated bits -> one word, bit b of word i = bit 32*i + b of the drop's bitstream
AX_HDN inline void ax_pack_item(const AxWave& w, int64_t wg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::edge_base, wg * 32);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int64_t j0 = wg * 32 - dr.edge_base;             // edge_base is a multiple of 64
    if (st.sm_status < 2 || j0 >= st.nbits_total) { w.bitw[wg] = 0; return; }
    const uint8_t* B = w.bit + dr.edge_base;
    uint32_t v = 0;
    for (int q = 0; q < 32; ++q) if (j0 + q < st.nbits_total && B[j0 + q]) v |= 1u << q;
    w.bitw[wg] = v;
}

AX_HD uint32_t ax_brev32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
// the 32 bits starting at bit position p of the drop's bitstream, first bit in the MSB (binListToHex order)
AX_HD uint32_t ax_frame_word(const uint32_t* bw, int64_t p) {
    const int64_t wi = p >> 5; const int sh = (int)(p & 31);
    const uint64_t two = (uint64_t)bw[wi] | ((uint64_t)bw[wi + 1] << 32);
    return ax_brev32((uint32_t)(two >> sh));
}

// parse.py:68 for every start position: '10' header, CRC, r7500 > 0  ->  one mask bit per position
AX_HDN inline void ax_valid_item(const AxWave& w, int64_t wg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::edge_base, wg * 32);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int64_t j0 = wg * 32 - dr.edge_base;
    if (st.sm_status < 2 || j0 >= st.nbits_total) { w.validw[wg] = 0; return; }
    const uint32_t* bw = w.bitw + dr.edge_base / 32;
    uint32_t v = 0;
    for (int q = 0; q < 32; ++q) {
        const int64_t p = j0 + q;
        if (p + 32 > st.nbits_total) break;
        const uint32_t wd = ax_frame_word(bw, p);
        if ((wd >> 30) == 2u && ax_crc_ok(wd) && ax_lvl7500(w, dr, p) > 0.0) v |= 1u << q;
    }
    w.validw[wg] = v;
}

// next position >= p whose mask bit is set (or `limit` if none below limit)
AX_HD int64_t ax_next_valid(const uint32_t* vw, int64_t p, int64_t limit) {
    while (p < limit) {
        const uint32_t v = vw[p >> 5] >> (p & 31);
        if (v) {
#ifdef __CUDA_ARCH__
            const int z = __ffs((int)v) - 1;
#else
            const int z = __builtin_ctz(v);
#endif
            p += z;
            return p < limit ? p : limit;
        }
        p = ((p >> 5) + 1) << 5;
    }
    return limit;
}

// ---- greedy frame synchronisation (parse.py:57-89 over AXCTDprocessor.py's per-iteration buffers) ----
// The scan of iteration k is a function of one integer, the bit position `cur` where the previous
// iteration stopped: profile trim (AXCTDprocessor.py:545-551), then candidate by candidate up to
// limit = NB - 32 (parse.py:57), consuming the scanned bits (:618-621).  ax_frames_scan is that
// function; fr != nullptr records the frame positions.
//   returns 0, or AXCTD_DROP_TRIM_INDEX / AXCTD_DROP_CAPACITY
AX_HD int ax_frames_scan(const AxWave& w, const AxDrop& dr, const AxState& st, const AxChunk& ch, int k, int64_t cur,
                         axctd_frame* fr, int32_t room, int32_t* cnt_out, int64_t* cur_out) {
    const int32_t* I = w.edge_idx + dr.edge_base;
    const uint32_t* vw = w.validw + dr.edge_base / 32;
    const int64_t NI = ch.edge_off + ch.n_edges, NB = ch.bit_off + ch.n_edges - 1;
    int32_t nf = 0;
    if (cur < NI && (int64_t)I[cur] <= ch.profstart) {                 // AXCTDprocessor.py:545-551 (profstartind of THIS iteration)
        const int64_t f = ax_first_gt(I, cur, NI, ch.profstart);
        if (f < 0) return AXCTD_DROP_TRIM_INDEX;
        cur = f;
    }
    const int64_t limit = NB - 32;                                     // parse.py:57 `while s < numbits - 32`
    int64_t p = cur;
    while (p < limit) {
        p = ax_next_valid(vw, p, limit);
        if (p >= limit) break;
        if (nf >= room) return AXCTD_DROP_CAPACITY;
        if (fr) { fr[nf].edge_index = p; fr[nf].chunk = k; }           // bit position; ax_calib_item resolves it
        ++nf;
        p += 32;
    }
    if (p > cur) cur = p;                                              // AXCTDprocessor.py:618-621
    *cnt_out = nf; *cur_out = cur;
    return 0;
}

// Speculative scan of one iteration: the scans lock onto the frame grid, so a scan started a few
// frames before the end of the previous iteration's bits arrives at this iteration with the same
// `cur` as the true one (checked, and repaired if not, by ax_frames_chain_item).
#define AX_FRAME_WARM 256
AX_HDN inline void ax_frames_spec_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.status != 0 || st.sm_status < 2 || st.k2 < 0 || st.nedges_total == 0 || k < st.k2 || k >= st.n_chunks || k >= dr.chunk_cap) return;
    AxChunk* chs = w.chunk + dr.chunk_base;
    AxChunk& ch = chs[k];
    ch.scan_from = -1; ch.scan_cnt = 0; ch.scan_end = -1;
    if (ch.n_edges <= 0) return;
    int kp = k - 1;
    while (kp >= st.k2 && chs[kp].n_edges <= 0) --kp;
    int64_t cur = 0;
    if (kp >= st.k2) {
        const AxChunk& pc = chs[kp];
        const int64_t plimit = pc.bit_off + pc.n_edges - 1 - 32;
        const uint32_t* vw = w.validw + dr.edge_base / 32;
        int64_t p = plimit - AX_FRAME_WARM;
        if (p < 0) p = 0;
        cur = p;
        while (p < plimit) {
            p = ax_next_valid(vw, p, plimit);
            if (p >= plimit) break;
            p += 32;
        }
        if (p > cur) cur = p;
    }
    int32_t cnt; int64_t cend;
    if (ax_frames_scan(w, dr, st, ch, k, cur, nullptr, 0x7fffffff, &cnt, &cend) != 0) return;       // left to the chain pass
    ch.scan_from = cur; ch.scan_cnt = cnt; ch.scan_end = cend;
}

// Sequential pass over the iterations of one drop: accept the speculative scans whose start matches,
// redo the others, assign the frame ranges.
AX_HDN inline void ax_frames_chain_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    st.n_frames = 0;
    if (st.status != 0 || st.sm_status < 2 || st.k2 < 0 || st.nedges_total == 0) return;
    AxChunk* ch = w.chunk + dr.chunk_base;
    int64_t cur = 0;
    int32_t nf = 0;
    for (int k = st.k2; k < st.n_chunks; ++k) {
        ch[k].frame_begin = nf; ch[k].frame_end = nf;
        if (ch[k].n_edges <= 0) continue;
        if (ch[k].scan_from != cur) {                                   // mis-speculated (or first) iteration
            int32_t cnt; int64_t cend;
            const int err = ax_frames_scan(w, dr, st, ch[k], k, cur, nullptr, 0x7fffffff, &cnt, &cend);
            if (err) { ax_raise(st, err, k); return; }
            ch[k].scan_from = cur; ch[k].scan_cnt = cnt; ch[k].scan_end = cend;
            st.n_frame_respec++;
        }
        if (nf + ch[k].scan_cnt > dr.frame_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; return; }
        nf += ch[k].scan_cnt;
        ch[k].frame_end = nf;
        cur = ch[k].scan_end;
    }
    st.n_frames = nf;
}

// Record the frame positions of one iteration (its true start is known now).
AX_HDN inline void ax_frames_write_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.status != 0 || st.n_frames == 0 || st.k2 < 0 || k < st.k2 || k >= st.n_chunks || k >= dr.chunk_cap) return;
    const AxChunk& ch = w.chunk[cg];
    if (ch.n_edges <= 0 || ch.frame_end <= ch.frame_begin) return;
    int32_t cnt; int64_t cend;
    ax_frames_scan(w, dr, st, ch, k, ch.scan_from, w.frame + dr.frame_base + ch.frame_begin, ch.frame_end - ch.frame_begin, &cnt, &cend);
}

// generic one-thread form of the three passes (reference order)
AX_HDN inline void ax_frames_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    for (int k = 0; k < dr.chunk_cap; ++k) ax_frames_spec_item(w, dr.chunk_base + k);
    ax_frames_chain_item(w, d);
    for (int k = 0; k < dr.chunk_cap; ++k) ax_frames_write_item(w, dr.chunk_base + k);
}

// ---- PSS-78 (gsw_sp_from_c of GSW-C; reference parse.py:132) ----------------
AX_HD double ax_sp_poly(double x, double ft) {
    return 0.0080 + (-0.1692 + (25.3851 + (14.0941 + (-7.0261 + 2.7081 * x) * x) * x) * x) * x
         + ft * (0.0005 + (-0.0056 + (-0.0066 + (-0.0375 + (0.0636 + -0.0144 * x) * x) * x) * x) * x);
}
AX_HD double ax_dsp_poly(double x, double ft) {
    return -0.1692 + (2 * 25.3851 + (3 * 14.0941 + (4 * -7.0261 + 5 * 2.7081 * x) * x) * x) * x
         + ft * (-0.0056 + (2 * -0.0066 + (3 * -0.0375 + (4 * 0.0636 + 5 * -0.0144 * x) * x) * x) * x);
}
AX_HD double ax_hill_ratio(double t) {
    const double g0 = 2.641463563366498e-1, g1 = 2.007883247811176e-4, g2 = -4.107694432853053e-6,
                 g3 = 8.401670882091225e-8, g4 = -1.711392021989210e-9, g5 = 3.374193893377380e-11,
                 g6 = -5.923731174730784e-13, g7 = 8.057771569962299e-15, g8 = -7.054313817447962e-17,
                 g9 = 2.859992717347235e-19;
    const double t68 = t * 1.00024;
    const double ft = (t68 - 15.0) / (1.0 + 0.0162 * (t68 - 15.0));
    const double rtx0 = g0 + t68 * (g1 + t68 * (g2 + t68 * (g3 + t68 * (g4 + t68 * (g5 + t68 * (g6 + t68 * (g7 + t68 * (g8 + t68 * g9))))))));
    double dsp = ax_dsp_poly(rtx0, ft);
    const double sp_est = ax_sp_poly(rtx0, ft);
    double rtx = rtx0 - (sp_est - 2.0) / dsp;
    const double rtxm = 0.5 * (rtx + rtx0);
    dsp = ax_dsp_poly(rtxm, ft);
    rtx = rtx0 - (sp_est - 2.0) / dsp;
    const double x = 400.0 * rtx * rtx, sq = 10.0 * rtx;
    const double part1 = 1.0 + x * (1.5 + x), part2 = 1.0 + sq * (1.0 + sq * (1.0 + sq));
    return 2.0 / (2.0 - 0.0080 / part1 - 0.0005 * ft / part2);
}
AX_HD double ax_sp_from_c(double C, double t, double p) {
    const double t68 = t * 1.00024;
    const double ft = (t68 - 15.0) / (1.0 + 0.0162 * (t68 - 15.0));
    const double r = 0.023302418791070513 * C;
    const double rt_lc = 0.6766097 + (2.00564e-2 + (1.104259e-4 + (-6.9698e-7 + 1.0031e-9 * t68) * t68) * t68) * t68;
    const double rp = 1.0 + (p * (2.070e-5 + -6.370e-10 * p + 3.989e-15 * p * p))
                          / (1.0 + 3.426e-2 * t68 + 4.464e-4 * t68 * t68 + (4.215e-1 + -3.107e-3 * t68) * r);
    double rt = r / (rp * rt_lc);
    if (rt < 0.0) rt = ax_nan();
    const double rtx = sqrt(rt);
    double sp = ax_sp_poly(rtx, ft);
    if (sp < 2.0) {
        const double x = 400.0 * rt, sq = 10.0 * rtx;
        const double part1 = 1.0 + x * (1.5 + x), part2 = 1.0 + sq * (1.0 + sq * (1.0 + sq));
        sp = ax_hill_ratio(t) * (sp - 0.0080 / part1 - 0.0005 * ft / part2);
    }
    if (sp < 0.0) sp = ax_nan();
    return sp;
}

// parse.py:297-301: output = 0; output += c[i] * x**i
AX_HD double ax_dataconvert(double x, const double* cf) {
    double out = cf[0];                                   // 0 + c0 * x**0
    out = ax_add(out, ax_mul(cf[1], x));
    const double x2 = ax_mul(x, x);
    out = ax_add(out, ax_mul(cf[2], x2));
    out = ax_add(out, ax_mul(cf[3], ax_mul(x2, x)));
    return out;
}
AX_HD double ax_round2(double v) { return ax_div(rint(ax_mul(v, 100.0)), 100.0); }   // np.round(v, 2)

AX_HDN inline void ax_calib_item(const AxWave& w, int64_t fg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::frame_base, fg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    if (fg - dr.frame_base >= st.n_frames || st.status != 0) return;
    const AxCfg& c = w.cfg[dr.cfg];
    axctd_frame& f = w.frame[fg];
    {   // the scan stored the bit position in edge_index: resolve word, PCM index, time and levels
        const int64_t p = f.edge_index;
        const int64_t ei = w.edge_idx[dr.edge_base + p];
        f.word = ax_frame_word(w.bitw + dr.edge_base / 32, p);
        f.edge_index = ei;
        f.time_raw = ax_div((double)(ei - w.chunk[dr.chunk_base + f.chunk].profstart), c.fs);     // AXCTDprocessor.py:554
        f.r400_raw = ax_lvl400(w, dr, p); f.r7500_raw = ax_lvl7500(w, dr, p);
        f.hex_returned = 0;
    }
    f.cint = (int32_t)((f.word >> 18) & 0xFFF);           // bits 2..13  (parse.py:107)
    f.tint = (int32_t)((f.word >> 6) & 0xFFF);            // bits 14..25 (parse.py:106)
    const double z = ax_dataconvert(f.time_raw, st.zc_used);                    // parse.py:117
    const double tun = (f.tint >= 0 && f.tint <= c.lut_len - 1) ? c.lut[f.tint] : ax_nan();   // parse.py:120-123
    const double cun = ax_div((double)(f.cint * 60), 4096.0);                   // parse.py:125
    const double T = ax_dataconvert(tun, st.tc_used);
    const double C = ax_dataconvert(cun, st.cc_used);
    const double S = ax_sp_from_c(C, T, z);
    f.depth_raw = z; f.temperature_raw = T; f.conductivity_raw = C; f.salinity_raw = S;
    f.time_s = ax_round2(ax_add(f.time_raw, st.firstpointtime));                // AXCTDprocessor.py:560
    f.depth = ax_round2(z); f.temperature = ax_round2(T); f.conductivity = ax_round2(C); f.salinity = ax_round2(S);
    f.r400 = ax_round2(f.r400_raw); f.r7500 = ax_round2(f.r7500_raw);
    const bool bad = f.r7500 < c.min_dr7500_inprof || f.r400 < c.min_r400_inprof || f.temperature < c.tlims[0]
                  || f.temperature > c.tlims[1] || f.salinity < c.slims[0] || f.salinity > c.slims[1];   // :572-574
    f.keep = bad ? 0 : 1;
}

// np.percentile(v, q) (method 'linear') on an ascending array without NaN
AX_HD double ax_percentile_sorted(const double* v, int n, double q) {
    const double virt = ax_mul((double)(n - 1), q);
    int64_t prev = (int64_t)floor(virt), next = prev + 1;
    if (virt >= (double)(n - 1)) { prev = n - 1; next = n - 1; }
    if (virt < 0) { prev = 0; next = 0; }
    const double gamma = ax_sub(virt, floor(virt));
    const double a = v[prev], b = v[next];
    const double diff = ax_sub(b, a);
    double r = ax_add(a, ax_mul(diff, gamma));
    if (gamma >= 0.5) r = ax_sub(b, ax_mul(diff, ax_sub(1.0, gamma)));
    return r;
}

AX_HD void ax_sort_small(double* v, int n) {
    for (int i = 1; i < n; ++i) { const double x = v[i]; int j = i - 1; while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; } v[j + 1] = x; }
}

// AXCTDprocessor.py:576-613 for the frames parsed in one iteration
AX_HDN inline void ax_qc_item(const AxWave& w, int64_t cg, double* scratch) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.status != 0 || st.k2 < 0 || k < st.k2 || k >= st.n_chunks) return;
    AxChunk& ch = w.chunk[cg];
    ch.n_rows = 0; ch.n_hex = 0;
    const int nfr = ch.frame_end - ch.frame_begin;
    if (nfr <= 0) return;
    axctd_frame* fr = w.frame + dr.frame_base + ch.frame_begin;
    double* tv = scratch + 2 * ((int64_t)dr.frame_base + ch.frame_begin);
    double* sv = tv + nfr;
    int n = 0; bool tnan = false, snan = false;
    for (int i = 0; i < nfr; ++i) if (fr[i].keep) {
        tv[n] = fr[i].temperature; sv[n] = fr[i].salinity;
        tnan |= isnan(tv[n]); snan |= isnan(sv[n]); ++n;
    }
    if (n == 0) return;
    double Tlo = ax_nan(), Thi = ax_nan(), Slo = ax_nan(), Shi = ax_nan();
    if (!tnan) {
        ax_sort_small(tv, n);
        const double m = ax_percentile_sorted(tv, n, 0.5);
        Tlo = ax_sub(m, ax_mul(10.0, ax_sub(m, ax_percentile_sorted(tv, n, 0.15))));
        Thi = ax_add(m, ax_mul(10.0, ax_sub(ax_percentile_sorted(tv, n, 0.85), m)));
    }
    if (!snan) {
        ax_sort_small(sv, n);
        const double m = ax_percentile_sorted(sv, n, 0.5);
        Slo = ax_sub(m, ax_mul(10.0, ax_sub(m, ax_percentile_sorted(sv, n, 0.15))));
        Shi = ax_add(m, ax_mul(10.0, ax_sub(ax_percentile_sorted(sv, n, 0.85), m)));
    }
    int rows = 0;
    for (int i = 0; i < nfr; ++i) if (fr[i].keep) {
        const double T = fr[i].temperature, S = fr[i].salinity;
        if (T < Tlo || T > Thi || S < Slo || S > Shi) fr[i].keep = 0; else ++rows;
    }
    ch.n_rows = rows;
    if (rows > 0) { ch.n_hex = nfr; for (int i = 0; i < nfr; ++i) fr[i].hex_returned = 1; }   // :611-612
}

// ---- results as they leave the device ---------------------------------------------------------
AX_HD int32_t ax_centi(double v, int32_t* flags, bool narrow) {    // v is already np.round(v, 2)
    if (isnan(v)) return narrow ? AXCTD_ROW_NAN16 : AXCTD_ROW_NAN;
    const double q = rint(ax_mul(v, 100.0));
    if (!(fabs(q) < (narrow ? 32767.0 : 2147483000.0)) || ax_div(q, 100.0) != v) { *flags |= AXCTD_ROW_WIDE; return 0; }
    return (int32_t)q;
}
AX_HDN inline void ax_row_item(const AxWave& w, int64_t fg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::frame_base, fg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    if (fg - dr.frame_base >= st.n_frames || st.status != 0) return;
    const axctd_frame& f = w.frame[fg];
    axctd_row r;
    r.word = f.word;
    int32_t fl = (f.keep ? AXCTD_ROW_KEEP : 0) | (f.hex_returned ? AXCTD_ROW_HEX : 0);
    r.time_c = ax_centi(f.time_s, &fl, false); r.depth_c = ax_centi(f.depth, &fl, false);
    r.temperature_c = (int16_t)ax_centi(f.temperature, &fl, true); r.conductivity_c = (int16_t)ax_centi(f.conductivity, &fl, true);
    r.salinity_c = (int16_t)ax_centi(f.salinity, &fl, true);
    r.r400_c = (int16_t)ax_centi(f.r400, &fl, true); r.r7500_c = (int16_t)ax_centi(f.r7500, &fl, true);
    r.flags = (uint16_t)fl;
    w.row[fg] = r;
}
AX_HDN inline void ax_chunkout_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (k >= st.n_chunks || k >= dr.chunk_cap) return;
    const AxChunk& c = w.chunk[cg];
    axctd_chunk o;
    o.s = c.s; o.e = c.e; o.status = c.status; o.n_power_total = c.pw_off + c.np;
    const bool dem = st.k0 >= 0 && k >= st.k0 && c.n_edges > 0;
    o.n_bits = dem ? c.n_edges - 1 : -1;
    o.first_edge = dem ? (int32_t)(c.first_edge - c.s) : -1; o.last_edge = dem ? (int32_t)(c.true_last - c.s) : -1;
    o.n_head_edges = dem ? c.n_head_edges : 0;
    o.n_rows = c.n_rows; o.n_hex = c.n_hex; o.scale = c.scale; o.profstartind = c.profstart;
    w.chunk_out[cg] = o;
}
