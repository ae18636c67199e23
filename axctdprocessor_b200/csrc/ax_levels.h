// ax_levels.h -- tone powers, detection state machine, bit assembly.
//
//   ax_plan0_item      fixed chunk grid while status == 0            AXCTDprocessor.py:293-304, 332-333
//   ax_plan_tones_item power_inds of the demodulated chunks          AXCTDprocessor.py:357
//   ax_tone_direct     |sum x e^{j theta m}| at 400/7500/dead Hz     AXCTDprocessor.py:358-364
//   ax_sm_item         smoothing, log ratios, pulse / tone detection AXCTDprocessor.py:367-408, demodulate.py:39-48
//   ax_emit_item       bit edges, per-edge signal levels             AXCTDprocessor.py:413-429
//   ax_scale_item      mark/space scale calibration                  AXCTDprocessor.py:459-468, demodulate.py:124-157
//   ax_bits_item       bit decisions and confidence                  demodulate.py:109-114
#pragma once
#include "ax_dsp.h"

AX_HD int32_t ax_grid_count(int64_t s, int64_t e, const AxCfg& c) {
    const int64_t span = e - c.n_power - s;              // len(range(s, e - N_power, d_pcm))
    return span > 0 ? (int32_t)((span + c.d_pcm - 1) / c.d_pcm) : 0;
}

AX_HDN inline void ax_plan0_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    AxChunk* ch = w.chunk + dr.chunk_base;
    int k = 0;
    int64_t s = 0;
    int32_t pc = 0;
    int par = 1;
    if (w.streaming) {
        // a growing recording: the grid is extended by the iterations that have become complete; once the pulse is
        // found the chunk chain owns the iterations from k0 on
        if (st.sm_status >= 1 || st.status != 0) return;
        k = st.n_fixed;
        if (k > 0) { s = ch[k - 1].e; pc = ch[k - 1].pw_off + ch[k - 1].np; par = st.par_levels; }
    }
    while (true) {
        if (w.streaming == 1) { if (s + c.chunk_len >= dr.n) break; }      // :299 would cut it short: not complete yet
        else if (dr.n - s < 4 * (int64_t)c.n_power) break;
        if (k >= dr.chunk_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; break; }
        int64_t e = s + c.chunk_len;
        if (e >= dr.n) e = dr.n - 1;
        AxChunk& q = ch[k];
        q.s = s; q.e = e; q.pw_off = pc; q.np = ax_grid_count(s, e, c);
        q.n_edges = 0; q.n_head_edges = 0; q.err = 0; q.status = 0; q.n_rows = 0; q.n_hex = 0;
        q.frame_begin = q.frame_end = 0; q.scale = c.scale0; q.mean7500 = ax_nan();
        q.spec_last = q.true_last = -1; q.g_first = -1; q.q_last = -1; q.bit_off = q.edge_off = 0; q.first_edge = -1;
        if (pc + q.np > dr.pw_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; break; }
        if (q.np < 10) par = 0;
        pc += q.np;
        ++k;
        s = e;
    }
    st.n_fixed = k;
    st.n_chunks = k;
    st.par_levels = par;
    st.searching = k > 0 ? 1 : 0;
    if (w.streaming && st.next_sm_chunk >= k) st.searching = 0;      // nothing new to look at in this run
}

// power_inds of one chunk (AXCTDprocessor.py:357)
AX_HDN inline void ax_pwfill_item(const AxWave& w, int64_t cg, int phase_b) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (k >= dr.chunk_cap || st.status >= AXCTD_DROP_CAPACITY) return;
    if (!phase_b) { if (k >= st.n_fixed) return; }
    else if (st.sm_status < 1 || k <= st.k0 || k >= st.n_chunks) return;
    const AxChunk& q = w.chunk[cg];
    const AxCfg& c = w.cfg[dr.cfg];
    for (int j = 0; j < q.np; ++j) w.pw_ind[dr.pw_base + q.pw_off + j] = q.s + (int64_t)j * c.d_pcm;
}

// After the chain is final: power grid of the chunks after the first demodulated one.
AX_HDN inline void ax_plan_tones_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (st.sm_status < 1) return;
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    int32_t pc = ch[st.k0].pw_off + ch[st.k0].np;
    for (int k = st.k0 + 1; k < st.n_chunks; ++k) {
        AxChunk& q = ch[k];
        q.pw_off = pc; q.np = ax_grid_count(q.s, q.e, c);
        if (k >= st.next_sm_chunk) {                      // (iterations the state machine has been through keep their records: streaming)
            q.status = 0; q.n_rows = 0; q.n_hex = 0; q.frame_begin = q.frame_end = 0; q.scale = c.scale0; q.mean7500 = ax_nan();
        }
        if (pc + q.np > dr.pw_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; q.np = 0; st.n_chunks = k; break; }
        if (q.np < 10) st.par_levels = 0;
        pc += q.np;
    }
}

// Is power sample `slot` part of the current tone launch?  (phase 0: detection round on the fixed grid,
// phase 1: the demodulated chunks after the first)
AX_HD bool ax_tone_slot_active(const AxWave& w, int64_t slot, int phase_b, int* d_out) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::pw_base, slot);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    *d_out = d;
    if (st.status >= AXCTD_DROP_CAPACITY) return false;
    const int32_t i = (int32_t)(slot - dr.pw_base);
    const AxChunk* ch = w.chunk + dr.chunk_base;
    int32_t lo, hi;
    if (!phase_b) {
        const int ka = w.pa_lo, kb = w.pa_hi < st.n_fixed ? w.pa_hi : st.n_fixed;
        if (!st.searching || kb <= ka) return false;
        lo = ch[ka].pw_off; hi = ch[kb - 1].pw_off + ch[kb - 1].np;
    }
    else {
        if (st.sm_status < 1 || st.n_chunks <= st.k0 + 1) return false;
        lo = ch[st.k0].pw_off + ch[st.k0].np; hi = ch[st.n_chunks - 1].pw_off + ch[st.n_chunks - 1].np;
        if (w.streaming && st.next_sm_chunk > st.k0 + 1) lo = st.next_sm_chunk < st.n_chunks ? ch[st.next_sm_chunk].pw_off : hi;
    }
    return i >= lo && i < hi;
}

// One power sample, straightforward evaluation of AXCTDprocessor.py:358-364 (option "tone_direct";
// the production path is ax_toneblock_item + ax_tonewin_item, which reads the PCM once).
AX_HDN inline void ax_tone_direct_item(const AxWave& w, int64_t slot, int phase_flags) {
    const int phase_b = phase_flags & 1;
    int d;
    if (!ax_tone_slot_active(w, slot, phase_b, &d)) return;
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const AxCfg& c = w.cfg[dr.cfg];
    if ((phase_flags & 2) && ax_tone_blocked_ok(c)) return;      // already done by the blocked path
    const AxSrc x = ax_src(w, dr);
    const int64_t c0 = w.pw_ind[slot];
    const double kmul = st.inv_ampl, kadd = -(st.dc * st.inv_ampl);
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int m = 0; m < c.n_power; ++m) {
        const double u = ax_fma(ax_get(x, c0 + m), kmul, kadd);
        const double* t6 = c.tone_cs + 6 * (int64_t)m;
        for (int q = 0; q < 6; ++q) a[q] = ax_fma(u, t6[q], a[q]);
    }
    w.pw_raw[0 * (int64_t)w.pw_total + slot] = hypot(a[0], a[1]);
    w.pw_raw[1 * (int64_t)w.pw_total + slot] = hypot(a[2], a[3]);
    w.pw_raw[2 * (int64_t)w.pw_total + slot] = hypot(a[4], a[5]);
}

// ---- tone powers from aligned block sums --------------------------------------------------------
// |sum_m u[c+m] e^{j theta m}| does not depend on the phase reference, so the window [c, c+N) is split at
// the AX_TB-aligned block boundaries: B_f[j] = sum_{m<AX_TB} x[AX_TB*j+m] e^{j theta_f m} is computed once
// per block of raw samples in the same pass that takes the mean and max|x| (no chunk start is known
// yet), and a window is  sum_j e^{j theta_f (AX_TB*j - c)} B_f[j]  plus its two ragged ends, normalised
// afterwards: sum_m (x-dc)/ampl e^{j theta m} = (S_x - dc * T_f) / ampl with T_f = sum_m e^{j theta_f m}.
AX_HDN inline void ax_toneblock_item(const AxWave& w, int64_t tbg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::tb_base, tbg);
    const AxDrop& dr = w.drop[d];
    const int64_t j = tbg - dr.tb_base;
    if (j >= dr.ntb || (w.only_xf && dr.xf_off < 0)) return;       // (k_stats_tones already did the int16 drops)
    if (w.streaming && j < w.st[d].tb_done) return;                // summed by an earlier run of the growing recording
    const AxCfg& c = w.cfg[dr.cfg];
    const AxSrc x = ax_src(w, dr);
    double a[6] = {0, 0, 0, 0, 0, 0};
    for (int m = 0; m < AX_TB; ++m) {
        const double xd = ax_get(x, j * AX_TB + m);
        const double* t6 = c.tone_cs + 6 * (int64_t)m;
        for (int q = 0; q < 6; ++q) a[q] = ax_fma(xd, t6[q], a[q]);
    }
    for (int q = 0; q < 6; ++q) w.tb_sum[tbg * 6 + q] = a[q];
}

// partial sums of one window over the terms i = lane, lane+nl, ... (ragged-end samples first, then blocks)
AX_HD void ax_tonewin_partial(const AxWave& w, const AxDrop& dr, const AxCfg& c, int64_t cstart, int lane, int nl, double* a) {
    const AxSrc x = ax_src(w, dr);
    const int np = c.n_power;
    const int64_t cend = cstart + np;
    int64_t j0 = (cstart + AX_TB - 1) / AX_TB, j1 = cend / AX_TB;       // full blocks j0 .. j1-1
    if (j1 > dr.ntb) j1 = dr.ntb;
    if (j1 < j0) j1 = j0;
    const int head_n = (int)(j0 * AX_TB - cstart);                       // samples before the first full block
    const int tail_off = (int)(j1 * AX_TB - cstart);                     // window offset of the first sample after the last full block
    const double* t0 = c.tone_soa; const double* t1 = t0 + np; const double* t2 = t1 + np;
    const double* t3 = t2 + np; const double* t4 = t3 + np; const double* t5 = t4 + np;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0;
    for (int m = lane; m < head_n; m += nl) {
        const double xd = ax_get(x, cstart + m);
        a0 = ax_fma(xd, t0[m], a0); a1 = ax_fma(xd, t1[m], a1); a2 = ax_fma(xd, t2[m], a2);
        a3 = ax_fma(xd, t3[m], a3); a4 = ax_fma(xd, t4[m], a4); a5 = ax_fma(xd, t5[m], a5);
    }
    for (int m = tail_off + lane; m < np; m += nl) {
        const double xd = ax_get(x, cstart + m);
        a0 = ax_fma(xd, t0[m], a0); a1 = ax_fma(xd, t1[m], a1); a2 = ax_fma(xd, t2[m], a2);
        a3 = ax_fma(xd, t3[m], a3); a4 = ax_fma(xd, t4[m], a4); a5 = ax_fma(xd, t5[m], a5);
    }
    a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3; a[4] = a4; a[5] = a5;
    // block j sits m = head_n + AX_TB*jj samples into the window: e^{j theta m} = e^{j theta head_n} * e^{j theta AX_TB jj}
    // (one broadcast row of the table and a small per-config table instead of a scattered table row per block)
    const int nblk = (int)(j1 - j0);
    const double* B0 = w.tb_sum + (dr.tb_base + j0) * 6;
    double eh[6];
    for (int q = 0; q < 6; ++q) eh[q] = c.tone_soa[(int64_t)q * np + head_n];
    for (int jj = lane; jj < nblk; jj += nl) {
        const double* B = B0 + 6 * jj;
        const double* R = c.tone_rot[jj];
        for (int f = 0; f < 3; ++f) {
            const double cr = ax_fma(eh[2 * f], R[2 * f], -(eh[2 * f + 1] * R[2 * f + 1]));
            const double sn = ax_fma(eh[2 * f], R[2 * f + 1], eh[2 * f + 1] * R[2 * f]);
            const double br = B[2 * f], bi = B[2 * f + 1];
            a[2 * f] = ax_fma(br, cr, ax_fma(-bi, sn, a[2 * f]));
            a[2 * f + 1] = ax_fma(br, sn, ax_fma(bi, cr, a[2 * f + 1]));
        }
    }
}
AX_HD void ax_tonewin_finish(const AxWave& w, const AxCfg& c, const AxState& st, int64_t slot, const double* a) {
    for (int f = 0; f < 3; ++f) {
        const double re = (a[2 * f] - st.dc * c.tone_tsum[2 * f]) * st.inv_ampl;
        const double im = (a[2 * f + 1] - st.dc * c.tone_tsum[2 * f + 1]) * st.inv_ampl;
        w.pw_raw[f * (int64_t)w.pw_total + slot] = hypot(re, im);
    }
}
// generic one-thread form (the CUDA build sums a window with a warp: k_tone_windows)
AX_HDN inline void ax_tonewin_item(const AxWave& w, int64_t slot, int phase_b) {
    int d;
    if (!ax_tone_slot_active(w, slot, phase_b, &d)) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    if (!ax_tone_blocked_ok(c)) return;
    double a[6];
    ax_tonewin_partial(w, dr, c, w.pw_ind[slot], 0, 1, a);
    ax_tonewin_finish(w, c, w.st[d], slot, a);
}

// numpy's pairwise summation for a contiguous float64 vector of n <= 128
// (np.sum inside np.nanmean, AXCTDprocessor.py:393)
AX_HD double ax_np_sum_small(const double* a, int n) {
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r = ax_add(r, a[i]); return r; }
    double r[8];
    for (int q = 0; q < 8; ++q) r[q] = a[q];
    int i = 8;
    for (; i < n - (n % 8); i += 8) for (int q = 0; q < 8; ++q) r[q] = ax_add(r[q], a[i + q]);
    double res = ax_add(ax_add(ax_add(r[0], r[1]), ax_add(r[2], r[3])), ax_add(ax_add(r[4], r[5]), ax_add(r[6], r[7])));
    for (; i < n; ++i) res = ax_add(res, a[i]);
    return res;
}

// chunk range [klo, khi) whose power samples a levels / tone launch covers
AX_HD bool ax_level_range(const AxWave& w, const AxState& st, int phase_b, int* klo, int* khi) {
    if (!phase_b) {
        *klo = w.pa_lo; *khi = w.pa_hi < st.n_fixed ? w.pa_hi : st.n_fixed;
        return st.searching && *khi > *klo;
    }
    *klo = st.k0 + 1; *khi = st.n_chunks;
    if (w.streaming && st.next_sm_chunk > *klo) *klo = st.next_sm_chunk;      // a growing recording: the new iterations only
    return st.sm_status >= 1 && *khi > *klo;
}

AX_HD double ax_nanmean_raw(const double* raw, int lo, int hi) {     // np.nanmean(raw[lo..hi]) (demodulate.py:44,46)
    double tot = 0.0; int cnt = 0;
    for (int j = lo; j <= hi; ++j) { const double v = raw[j]; if (!isnan(v)) { tot = ax_add(tot, v); ++cnt; } }
    return cnt ? ax_div(tot, (double)cnt) : ax_nan();
}

// demodulate.py:39-48 + AXCTDprocessor.py:370-371 for ONE power sample.  The lagging box filter
// re-reads entries smoothed by earlier calls only in the first five samples of a chunk; those
// entries are themselves plain window means of raw values whenever the previous chunk holds at
// least ten samples (st.par_levels), so every sample can be evaluated independently.
AX_HDN inline void ax_levels_item(const AxWave& w, int64_t slot, int phase_b) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::pw_base, slot);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    if (!st.par_levels || st.status >= AXCTD_DROP_CAPACITY) return;
    int klo, khi;
    if (!ax_level_range(w, st, phase_b, &klo, &khi)) return;
    const AxChunk* ch = w.chunk + dr.chunk_base;
    const int i = (int)(slot - dr.pw_base);
    if (i < ch[klo].pw_off || i >= ch[khi - 1].pw_off + ch[khi - 1].np) return;
    int lo = klo, hi = khi - 1;                          // last chunk with pw_off <= i
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ch[mid].pw_off <= i) lo = mid; else hi = mid - 1; }
    const int pstart = ch[lo].pw_off;
    const int64_t PT = w.pw_total;
    double smv[3];
    for (int f = 0; f < 3; ++f) {
        const double* raw = w.pw_raw + f * PT + dr.pw_base;
        const int wl = (i < 5) ? 0 : i - 5;
        double tot = 0.0; int cnt = 0;
        for (int j = wl; j <= i; ++j) {
            const double v = (j >= pstart) ? raw[j] : ax_nanmean_raw(raw, j < 5 ? 0 : j - 5, j);
            if (!isnan(v)) { tot = ax_add(tot, v); ++cnt; }
        }
        smv[f] = cnt ? ax_div(tot, (double)cnt) : ax_nan();
        w.pw_sm[f * PT + slot] = smv[f];
    }
    w.r400[slot] = log10(ax_div(smv[0], smv[2]));
    w.r7500[slot] = log10(ax_div(smv[1], smv[2]));
}

// Sequential per-drop state machine over run() iterations.  phase 0: fixed grid
// until the 400 Hz pulse is found (in rounds of chunks [pa_lo, pa_hi)); phase 1: the demodulated
// chunks.  Smoothing is done here only for drops that cannot use ax_levels_item.
AX_HDN inline void ax_sm_item(const AxWave& w, int64_t d, int phase_b) {
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    if (st.status != 0 && !phase_b) return;
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int64_t PT = w.pw_total;
    double* raw[3]; double* sm[3];
    for (int f = 0; f < 3; ++f) { raw[f] = w.pw_raw + f * PT + dr.pw_base; sm[f] = w.pw_sm + f * PT + dr.pw_base; }
    double* r400 = w.r400 + dr.pw_base; double* r7500 = w.r7500 + dr.pw_base;
    const int64_t* pind = w.pw_ind + dr.pw_base;
    int klo, kend;
    if (!ax_level_range(w, st, phase_b, &klo, &kend)) return;
    const bool do_smooth = !st.par_levels;
    for (int k = st.next_sm_chunk; k < kend; ++k) {
        AxChunk& q = ch[k];
        const int pstart = q.pw_off, np = q.np;
        if (do_smooth) {
            // demodulate.py:39-48 called with startind = pstart (AXCTDprocessor.py:367-369)
            for (int f = 0; f < 3; ++f) {
                for (int i = pstart; i < pstart + np; ++i) {
                    const int lo = (i < 5) ? 0 : i - 5;
                    double tot = 0.0; int cnt = 0;
                    for (int jj = lo; jj <= i; ++jj) {
                        const double v = (jj < pstart) ? sm[f][jj] : raw[f][jj];
                        if (!isnan(v)) { tot = ax_add(tot, v); ++cnt; }
                    }
                    sm[f][i] = cnt ? ax_div(tot, (double)cnt) : ax_nan();
                }
            }
            for (int i = pstart; i < pstart + np; ++i) {                    // :370-371
                r400[i] = log10(ax_div(sm[0][i], sm[2][i]));
                r7500[i] = log10(ax_div(sm[1][i], sm[2][i]));
            }
        }
        st.pcount = pstart + np;
        if (st.sm_status == 0) {                                        // :375-380
            for (int i = pstart; i < pstart + np; ++i) {
                if (r400[i] >= c.min_r400) { st.firstpulse400 = pind[i]; st.sm_status = 1; st.k0 = k; break; }
            }
        }
        if (st.sm_status >= 1 && st.pcount > 0) {
            const int64_t last_ind = pind[st.pcount - 1];
            const int64_t fp = st.firstpulse400;
            if (last_ind >= fp + c.off_5p5 && isnan(st.mean7500)) {     // :388-393
                int s75 = 0, e75 = 0; int64_t b1 = 0, b2 = 0;
                for (int jj = 0; jj < st.pcount; ++jj) {
                    int64_t d1 = fp + c.off_4p5 - pind[jj]; if (d1 < 0) d1 = -d1;
                    int64_t d2 = fp + c.off_5p5 - pind[jj]; if (d2 < 0) d2 = -d2;
                    if (jj == 0 || d1 < b1) { b1 = d1; s75 = jj; }
                    if (jj == 0 || d2 < b2) { b2 = d2; e75 = jj; }
                }
                double buf[128]; int nb = 0, cnt = 0;
                for (int jj = s75; jj < e75 && nb < 128; ++jj) { const double v = r7500[jj]; if (isnan(v)) buf[nb++] = 0.0; else { buf[nb++] = v; ++cnt; } }
                st.mean7500 = cnt ? ax_div(ax_np_sum_small(buf, nb), (double)cnt) : ax_nan();
                st.km = k;
            }
            if (last_ind > fp + c.off_trig_from) {                      // :397-408
                if (!isnan(st.mean7500) && st.sm_status == 1) {
                    for (int i = pstart; i < pstart + np; ++i) {
                        if (ax_sub(r7500[i], st.mean7500) >= c.min_dr7500) { st.profstartind = pind[i]; st.sm_status = 2; st.k2 = k; break; }
                    }
                } else if (c.trig_to > 0 && last_ind >= fp + c.off_trig_to) {
                    st.profstartind = fp + c.off_trig_to;
                    if (st.sm_status != 2) st.k2 = k;
                    st.sm_status = 2;
                }
                if (st.profstartind > 0 && st.firstpointtime <= 0) st.firstpointtime = ax_div((double)st.profstartind, c.fs);
            }
        }
        q.mean7500 = st.mean7500;
        q.status = st.sm_status;
        q.profstart = st.profstartind;
        st.next_sm_chunk = k + 1;
        if (!phase_b && st.sm_status >= 1) break;      // hand over to the chunk chain
        if (phase_b && !do_smooth && st.sm_status == 2 && !isnan(st.mean7500) && !(c.trig_to > 0)) {
            // nothing can change any more (status 2 is final unless the timeout branch of :404 is armed):
            // the remaining iterations only record the state
            for (int k2 = k + 1; k2 < kend; ++k2) { ch[k2].mean7500 = st.mean7500; ch[k2].status = 2; ch[k2].profstart = st.profstartind; }
            st.pcount = ch[kend - 1].pw_off + ch[kend - 1].np;
            st.next_sm_chunk = kend;
            break;
        }
    }
    if (!phase_b) {
        if (st.sm_status >= 1) { st.chain_from = st.k0; st.chain_end = 0; st.searching = 0; }
        else if (st.next_sm_chunk >= st.n_fixed) { st.n_chunks = st.n_fixed; st.chain_end = 1; st.searching = 0; }
        else w.flags[AX_FLAG_MORE] = 1;                // another detection round is needed
    }
}

// bit / edge array offsets of every demodulated chunk
AX_HDN inline void ax_offsets_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    st.nbits_total = 0; st.nedges_total = 0;
    if (st.sm_status < 1) return;
    AxChunk* ch = w.chunk + dr.chunk_base;
    int64_t nb = 0, ne = 0;
    for (int k = st.k0; k < st.n_chunks; ++k) {
        ch[k].bit_off = nb; ch[k].edge_off = ne;
        if (ch[k].n_edges > 0) { nb += ch[k].n_edges - 1; ne += ch[k].n_edges; }
    }
    if (ne > dr.edge_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, -1); w.flags[AX_FLAG_CAP] = 1; nb = 0; ne = 0; }
    st.nbits_total = nb; st.nedges_total = ne;
}

// AXCTDprocessor.py:413-429: bit edges of one chunk and the signal level of the
// nearest power sample of THIS chunk for every edge.
// One edge: t < n_head_edges comes from the exact head, otherwise `pos` is its dense crossing ordinal.
// edge t of iteration k from a crossing whose record (sample index, |S1|, |S2|) the caller has already fetched
AX_HD void ax_emit_edge_vals(const AxWave& w, const AxDrop& dr, AxState& st, const AxCfg& c, AxChunk& ch, int64_t cg, int k, int t,
                             int64_t idx, double v1, double v2, bool fused);
AX_HD void ax_emit_edge(const AxWave& w, const AxDrop& dr, AxState& st, const AxCfg& c, AxChunk& ch, int64_t cg, int k, int t, int64_t pos, bool fused = false) {
    int64_t idx; double v1, v2;
    if (t < ch.n_head_edges) {
        idx = ch.s + w.head_idx[cg * (int64_t)w.head_zc_cap_max + t];
        v1 = w.head_a1[cg * (int64_t)w.head_zc_cap_max + t]; v2 = w.head_a2[cg * (int64_t)w.head_zc_cap_max + t];
    } else {
        idx = w.zc_idx[dr.zc_base + pos]; v1 = w.zc_a1[dr.zc_base + pos]; v2 = w.zc_a2[dr.zc_base + pos];
    }
    ax_emit_edge_vals(w, dr, st, c, ch, cg, k, t, idx, v1, v2, fused);
}
AX_HD void ax_emit_edge_vals(const AxWave& w, const AxDrop& dr, AxState& st, const AxCfg& c, AxChunk& ch, int64_t cg, int k, int t,
                             int64_t idx, double v1, double v2, bool fused) {
    if (t >= ch.n_head_edges && t < ch.n_edges - 1 && idx + c.inset + c.npcm > ch.e) ax_raise(st, AXCTD_DROP_SHORT_WINDOW, k);   // demodulate.py:100-101
    const int64_t eo = dr.edge_base + ch.edge_off + t;
    w.edge_idx[eo] = (int32_t)idx;
    if (t < ch.n_edges - 1) {
        const int64_t slot = dr.edge_base + ch.bit_off + t;
        if (!fused) { w.a1[slot] = v1; w.a2[slot] = v2; }
        else {
            // the decision of ax_bits_decide right here (demodulate.py:109-114 with this iteration's scale); a bit whose
            // two tones are within bit_tol of each other is listed for a double-precision window (k_bits_recheck), and only
            // such bits keep their magnitudes
            const double p2 = ax_mul(v2, ch.scale);
            const double m = v1 > p2 ? v1 : p2;
            w.bit[slot] = (v1 >= p2) ? 1 : 0;
            if (fabs(v1 - p2) <= w.bit_tol * m) {
                w.a1[slot] = v1; w.a2[slot] = v2;
#if defined(__CUDA_ARCH__)
                const int pos_l = atomicAdd(&w.flags[AX_FLAG_FIXCNT], 1);
                if (pos_l < w.fix_cap) w.fix_list[pos_l] = slot; else w.flags[AX_FLAG_FIXOVF] = 1;
#endif
            }
        }
    }
    // np.argmin(np.abs(recent_pwrinds - ci)): nearest grid point, first on ties (:425,:428)
    if (ch.np > 0) {
        const int64_t off = idx - ch.s;
        int64_t jj = off / c.d_pcm;
        const int64_t rem = off - jj * c.d_pcm;
        if (2 * rem > c.d_pcm) ++jj;
        if (jj > ch.np - 1) jj = ch.np - 1;
        if (jj < 0) jj = 0;
        w.lvl_slot[eo] = (int32_t)(ch.pw_off + jj);
    } else w.lvl_slot[eo] = -1;
}
// The levels an edge takes (AXCTDprocessor.py:425-429): r400 and r7500 - mean7500pwr of its nearest power sample.
// Edges store the sample's index only; r7500m holds the difference per power sample (ax_emit_levels).
AX_HD double ax_lvl400(const AxWave& w, const AxDrop& dr, int64_t e) {
    const int32_t sl = w.lvl_slot[dr.edge_base + e];
    return sl < 0 ? ax_nan() : w.r400[dr.pw_base + sl];
}
AX_HD double ax_lvl7500(const AxWave& w, const AxDrop& dr, int64_t e) {
    const int32_t sl = w.lvl_slot[dr.edge_base + e];
    return sl < 0 ? ax_nan() : w.r7500m[dr.pw_base + sl];
}
AX_HD void ax_emit_levels(const AxWave& w, const AxDrop& dr, const AxChunk& ch, int first, int step) {
    for (int j = first; j < ch.np; j += step)
        w.r7500m[dr.pw_base + ch.pw_off + j] = ax_sub(w.r7500[dr.pw_base + ch.pw_off + j], ch.mean7500);
}

// position of the (r+1)-th set bit of m (r < popcount(m))
AX_HD int ax_select64(uint64_t m, int r) {
    for (int q = 0; q < r; ++q) m &= m - 1ull;
    return ax_ctz64(m);
}

// dense ordinal of edge t >= n_head_edges + n_pre of a chunk: the (t - n_head_edges - n_pre)-th canonical crossing from merge_pos
AX_HD int64_t ax_emit_canon_pos(const AxWave& w, const AxDrop& dr, const AxChunk& ch, int t) {
    const uint64_t* cmask = w.cmask + dr.tile_base;
    const int32_t* crank = w.crank + dr.tile_base;
    const int64_t r = ax_canon_rank(cmask, crank, ch.merge_pos) + (t - ch.n_head_edges - ch.n_pre);
    int64_t lo = ch.merge_pos / AX_TILE, hi = (ch.q_last) / AX_TILE;       // last tile with crank <= r
    while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if ((int64_t)crank[mid] <= r) lo = mid; else hi = mid - 1; }
    return lo * AX_TILE + ax_select64(cmask[lo], (int)(r - crank[lo]));
}

AX_HD bool ax_emit_active(const AxWave& w, const AxDrop& dr, const AxState& st, int k) {
    if (w.streaming && k < st.k_done) return false;      // edges of an earlier run of the growing recording: final
    return !(st.sm_status < 1 || k < st.k0 || k >= st.n_chunks || k >= dr.chunk_cap || st.nedges_total == 0);
}

// generic form: one thread per chunk (the CUDA build uses k_emit_chunk, one thread per edge)
AX_HDN inline void ax_emit_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (!ax_emit_active(w, dr, st, k)) return;
    AxChunk& ch = w.chunk[cg];
    if (ch.n_edges <= 0) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const uint8_t* nx = w.zc_nx + dr.zc_base;
    ax_emit_levels(w, dr, ch, 0, 1);
    int64_t pos = ch.g_first;
    for (int t = 0; t < ch.n_edges; ++t) {
        if (t >= ch.n_head_edges + ch.n_pre && ch.merge_pos >= 0) pos = ax_emit_canon_pos(w, dr, ch, t);
        ax_emit_edge(w, dr, st, c, ch, cg, k, t, pos);
        if (t >= ch.n_head_edges && t < ch.n_head_edges + ch.n_pre - 1) pos += nx[pos];
    }
}

// first index j in [0,n) with I[j] >= v (-1 if none); last index with I[j] <= v (-1 if none).
// I is only piecewise sorted (chunk joins overlap by a sample or two), so scan.
AX_HD int64_t ax_first_ge(const int32_t* I, int64_t n, int64_t v) {
    int64_t lo = ax_lower_bound(I, n, v);       // good guess; then make it exact
    int64_t j = lo - 64; if (j < 0) j = 0;
    for (; j < n; ++j) if ((int64_t)I[j] >= v) return j;
    return -1;
}
AX_HD int64_t ax_last_le(const int32_t* I, int64_t n, int64_t v) {
    int64_t hi = ax_upper_bound(I, n, v) + 64;
    if (hi > n) hi = n;
    for (int64_t j = hi - 1; j >= 0; --j) if ((int64_t)I[j] <= v) return j;
    return -1;
}
AX_HD int64_t ax_first_gt(const int32_t* I, int64_t from, int64_t n, int64_t v) {
    int64_t j = from + ax_upper_bound(I + from, n - from, v) - 64;     // good guess; then make it exact
    if (j < from) j = from;
    for (; j < n; ++j) if ((int64_t)I[j] > v) return j;
    return -1;
}

// ---- mark/space scale calibration: AXCTDprocessor.py:459-468 + demodulate.py:124-157 ----------
// Which iteration reads header 1, and which bit positions [a, hi) feed the histogram.
//   returns 1 found, 0 not (yet) available, <0 = -AXCTD_DROP_* error (chunk in *k_out)
AX_HD int ax_scale_find(const AxWave& w, int d, int* k_out, int64_t* a_out, int64_t* hi_out) {
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    const AxState& st = w.st[d];
    const AxChunk* ch = w.chunk + dr.chunk_base;
    const int32_t* I = w.edge_idx + dr.edge_base;
    const int64_t firstbin = I[0];
    const int64_t p1s = st.firstpulse400 + c.h1s, p1e = st.firstpulse400 + c.h1e;
    const int klast = (st.k2 >= 0) ? st.k2 : st.n_chunks - 1;
    for (int k = st.k0; k <= klast && k < st.n_chunks; ++k) {
        if (ch[k].n_edges <= 0) continue;
        const int64_t ni = ch[k].edge_off + ch[k].n_edges, nb = ch[k].bit_off + ch[k].n_edges - 1;
        const int64_t lastbin = I[ni - 1];
        if (!(firstbin <= p1s && lastbin >= p1e)) continue;
        const int64_t a = ax_first_ge(I, ni, p1s - c.half), b = ax_last_le(I, ni, p1e + c.half);
        *k_out = k;
        if (a < 0 || b < 0) return -AXCTD_DROP_TRIM_INDEX;
        *a_out = a; *hi_out = b < nb ? b : nb;
        return 1;
    }
    return 0;
}
// np.histogram bin of v for edges ed[0..nbins] (last bin closed), -1 if outside
AX_HD int ax_scale_bin(const AxCfg& c, double v) {
    const int nbins = c.n_hist_edges - 1;
    const double* ed = c.hist_edges;
    if (!(v >= ed[0]) || v > ed[nbins]) return -1;
    int lo = 0, up = nbins + 1;                                          // first edge > v
    while (lo < up) { int mid = (lo + up) >> 1; if (ed[mid] <= v) lo = mid + 1; else up = mid; }
    int bin = lo - 1;
    if (bin >= nbins) bin = nbins - 1;
    return bin;
}
// demodulate.py:136-153: cumulative percentage, centred slope, flattest stretch inside 30..65 %
AX_HD bool ax_scale_threshold(const AxCfg& c, const int* hist, int64_t npts, double* thr) {
    const int nbins = c.n_hist_edges - 1;
    double best = 0; int first = -1, last = -1;
    int64_t cs = 0;
    double cp[512];
    for (int q = 0; q < nbins; ++q) { cs += hist[q]; cp[q] = ax_div((double)(100 * cs), (double)npts); }
    const double* ctr = c.hist_centers;
    for (int q = 0; q < nbins; ++q) {
        if (!(cp[q] >= 30.0 && cp[q] <= 65.0)) continue;
        double sl;
        if (q == 0) sl = ax_div(ax_sub(cp[1], cp[0]), ax_sub(ctr[1], ctr[0]));
        else if (q == nbins - 1) sl = ax_div(ax_sub(cp[q], cp[q - 1]), ax_sub(ctr[q], ctr[q - 1]));
        else sl = ax_div(ax_sub(cp[q + 1], cp[q - 1]), ax_sub(ctr[q + 1], ctr[q - 1]));
        if (first < 0 || sl < best) { best = sl; first = q; last = q; }
        else if (sl == best) last = q;
    }
    if (first < 0) return false;
    *thr = ax_div(ax_add(ctr[first], ctr[last]), 2.0);
    return true;
}
// streaming: the calibration was made in an iteration that an earlier run already closed
AX_HD bool ax_scale_is_final(const AxWave& w, const AxState& st) {
    return w.streaming && st.header_read[0] && st.k1 >= 0 && st.k1 < st.k_done;
}
AX_HD void ax_scale_reset(const AxWave& w, int d) {
    AxState& st = w.st[d];
    st.scale = w.cfg[w.drop[d].cfg].scale0; st.k1 = -1; st.header_read[0] = 0; st.header_chunk[0] = -1; st.scale_switch_bit = st.nbits_total;
}
// per-iteration scale and the first bit decided with the calibrated one; chunks k = first, first+step, ...
AX_HD void ax_scale_spread(const AxWave& w, int d, int first, int step) {
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const double s0 = w.cfg[dr.cfg].scale0;
    for (int k = st.k0 + first; k < st.n_chunks; k += step) ch[k].scale = (st.k1 >= 0 && k > st.k1) ? st.scale : s0;
    if (first == 0) st.scale_switch_bit = (st.k1 >= 0 && st.k1 + 1 < st.n_chunks) ? ch[st.k1 + 1].bit_off : st.nbits_total;
}

// generic one-thread form (the CUDA build fills the histogram with a CTA: k_scale_block)
AX_HDN inline void ax_scale_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    if (ax_scale_is_final(w, st)) { ax_scale_spread(w, (int)d, 0, 1); return; }
    ax_scale_reset(w, (int)d);
    if (st.sm_status < 1 || st.nedges_total == 0) return;
    int k = -1; int64_t a = 0, hi = 0;
    const int found = ax_scale_find(w, (int)d, &k, &a, &hi);
    if (found < 0) { ax_raise(st, -found, k); return; }
    if (found) {
        const int nbins = c.n_hist_edges - 1;
        int hist[512];
        if (nbins > 512) { ax_raise(st, AXCTD_DROP_CAPACITY, k); return; }
        for (int q = 0; q < nbins; ++q) hist[q] = 0;
        const double* a1 = w.a1 + dr.edge_base; const double* a2 = w.a2 + dr.edge_base;
        for (int64_t jj = a; jj < hi; ++jj) {
            const int bin = ax_scale_bin(c, ax_div(ax_mul(a2[jj], c.scale0), a1[jj]));          // demodulate.py:102,110
            if (bin >= 0) hist[bin]++;
        }
        double thr;
        if (!ax_scale_threshold(c, hist, hi > a ? hi - a : 0, &thr)) { ax_raise(st, AXCTD_DROP_SCALE_EMPTY, k); return; }
        st.scale = ax_div(c.scale0, thr);
        st.k1 = k; st.header_read[0] = 1; st.header_chunk[0] = k;
    }
    ax_scale_spread(w, (int)d, 0, 1);
}

// ---- bit decisions (demodulate.py:102,109-114) ------------------------------------------------
// The mark / space magnitudes of the continuous pass carry fp32 accuracy (ax_window32).  Two
// consumers make discrete decisions from them: the scale calibration bins conf into 0.01-wide
// histogram bins (phase 0, before ax_scale_item) and the bit decision compares p1 with p2
// (phase 1).  Whenever a value lies within the tolerance of such a boundary, the window is
// re-evaluated in double precision from the int16 samples (ax_gwin_*) and replaced.

// run() iteration that demodulated bit j of the drop
AX_HD int ax_chunk_of_bit(const AxChunk* ch, int k0, int n_chunks, int64_t j) {
    int lo = k0, hi = n_chunks - 1;                      // last chunk with bit_off <= j
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ch[mid].bit_off <= j) lo = mid; else hi = mid - 1; }
    return lo;
}

struct AxBitFix { int32_t d; int64_t i, q0; };

// Does bit `slot` (bit j of drop d, demodulated by iteration k) need a double-precision window in this phase?
AX_HD bool ax_bits_need(const AxWave& w, int d, int k, int64_t slot, int phase, AxBitFix* fx) {
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int64_t j = slot - dr.edge_base;
    const AxCfg& c = w.cfg[dr.cfg];
    const double p1 = w.a1[slot];
    bool need = w.bitfix_all != 0;
    int64_t ei = -1;
    const AxChunk& ch = w.chunk[dr.chunk_base + k];
    const int64_t e = ch.edge_off + (j - ch.bit_off);
    if (phase == 0) {
        ei = w.edge_idx[dr.edge_base + e];
        // header-1 calibration bits (AXCTDprocessor.py:459-466).  The reference pairs bit j with edge-list
        // position j, which lags the bit's own edge by one entry per earlier iteration: keep a margin.
        const int64_t mg = (int64_t)(64.0 * c.fs / c.bitrate);
        const int64_t p1s = st.firstpulse400 + c.h1s - c.half - mg, p1e = st.firstpulse400 + c.h1e + c.half + mg;
        if (ei < p1s || ei > p1e) return false;
        if (!need) {
            const double v = ax_div(ax_mul(w.a2[slot], c.scale0), p1);
            const int nbins = c.n_hist_edges - 1;
            const double* ed = c.hist_edges;
            if (v <= ed[nbins] * 1.001) {
                int lo = 0, up = nbins + 1;                  // first edge > v
                while (lo < up) { const int mid = (lo + up) >> 1; if (ed[mid] <= v) lo = mid + 1; else up = mid; }
                // conf = a2*s/a1 with both magnitudes good to hist_tol of the stronger one
                const double a2v = w.a2[slot];
                const double tol = w.hist_tol * ((p1 > a2v ? p1 : a2v) / p1) * (c.scale0 + v);
                if (lo <= nbins && ed[lo] - v <= tol) need = true;
                if (lo >= 1 && v - ed[lo - 1] <= tol) need = true;
            }
        }
    } else if (!need) {
        const double scale = (j >= st.scale_switch_bit) ? st.scale : c.scale0;
        const double p2 = ax_mul(w.a2[slot], scale);
        const double m = p1 > p2 ? p1 : p2;
        need = fabs(p1 - p2) <= w.bit_tol * m;              // false for NaN
    }
    if (!need) return false;
    if (ei < 0) ei = w.edge_idx[dr.edge_base + e];
    fx->d = d; fx->i = ei; fx->q0 = ch.s;
    return true;
}

// Replace the magnitudes of bit `slot` by the double-precision ones (acc = summed ax_gwin_partial).
AX_HD void ax_bits_fix(const AxWave& w, int64_t slot, const AxBitFix& fx, const double* acc) {
    const AxDrop& dr = w.drop[fx.d];
    AxState& st = w.st[fx.d];
    double e1, e2;
    ax_gwin_finish(acc, fx.i, fx.q0, w.cfg[dr.cfg], st, &e1, &e2);
    const double o1 = w.a1[slot], o2 = w.a2[slot];
    // error of the fp32 magnitudes relative to the stronger of the two tones
    double rel = fabs(o1 - e1) > fabs(o2 - e2) ? fabs(o1 - e1) : fabs(o2 - e2);
    rel /= (e1 > e2 ? e1 : e2);
    if (rel == rel) {
        const float rf = (float)rel;
        int32_t bits;
        memcpy(&bits, &rf, sizeof(bits));
        AX_ATOMIC_MAX32(&st.err32_bits, bits);
    }
    AX_ATOMIC_ADD32(&st.n_recheck, 1);
    w.a1[slot] = e1; w.a2[slot] = e2;
}

AX_HD void ax_bits_decide(const AxWave& w, int d, int64_t slot) {
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int64_t j = slot - dr.edge_base;
    const double scale = (j >= st.scale_switch_bit) ? st.scale : w.cfg[dr.cfg].scale0;
    const double p1 = w.a1[slot];
    const double p2 = ax_mul(w.a2[slot], scale);
    // (the confidence p2 / p1 of demodulate.py:110 is not stored: axctd_batch_bits forms it from a1, a2 and the scale)
    w.bit[slot] = (p1 >= p2) ? 1 : 0;
}

// one thread does everything (generic form; the CUDA build shares the window sum across a warp)
AX_HDN inline void ax_bits_item(const AxWave& w, int64_t slot, int phase) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::edge_base, slot);
    const AxDrop& dr0 = w.drop[d];
    const AxState& st = w.st[d];
    const int64_t j = slot - dr0.edge_base;
    if (j >= st.nbits_total || st.sm_status < 1) return;
    const int k = ax_chunk_of_bit(w.chunk + dr0.chunk_base, st.k0, st.n_chunks, j);
    if (w.streaming && (k < st.k_done) && (phase == 1 || ax_scale_is_final(w, st))) return;     // decided by an earlier run
    AxBitFix fx;
    if (ax_bits_need(w, d, k, slot, phase, &fx)) {
        const AxDrop& dr = w.drop[fx.d];
        double acc[4];
        ax_gwin_partial(ax_src(w, dr), fx.i, fx.q0, w.cfg[dr.cfg], 0, 1, acc);
        ax_bits_fix(w, slot, fx, acc);
    }
    if (phase == 1) ax_bits_decide(w, d, slot);
}
