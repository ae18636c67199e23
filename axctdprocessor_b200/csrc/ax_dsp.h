// ax_dsp.h -- per-work-item bodies of the demodulation kernels.
//
// Every function here is the body of one CUDA thread (one "item").  The .cu
// file wraps them in __global__ launchers; the test-only emulation build runs
// the same bodies in host loops.
//
// Stage map (reference file:line):
//   ax_stats_item / ax_stats_fin   mean and max|x| of the int16 PCM     AXCTDprocessor.py:55-57
//   ax_filter_segment              Butterworth SOS low/band-pass, zero crossings and the
//                                  per-bit mark/space single-bin DFTs   demodulate.py:74-79, 99-102
//   ax_tiles_item / ax_walk_end    greedy bit-edge walk                 demodulate.py:85-93
//   ax_chain_item                  chunk chain s_k                      AXCTDprocessor.py:293-333
//   ax_head_item                   exact zero-state restart per chunk   demodulate.py:74 (sosfilt zero state)
#pragma once
#include "ax_types.h"

// ------------------------------------------------------------------ atomics
#if defined(__CUDA_ARCH__)
#define AX_ATOMIC_ADD64(p, v) atomicAdd((unsigned long long*)(p), (unsigned long long)(v))
#define AX_ATOMIC_ADD32(p, v) atomicAdd((int*)(p), (int)(v))
#define AX_ATOMIC_MAX32(p, v) atomicMax((int*)(p), (int)(v))
#else
#define AX_ATOMIC_ADD64(p, v) __atomic_fetch_add((long long*)(p), (long long)(v), __ATOMIC_RELAXED)
#define AX_ATOMIC_ADD32(p, v) __atomic_fetch_add((int*)(p), (int)(v), __ATOMIC_RELAXED)
static inline void ax_host_atomic_max(int* p, int v) {
    int cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}
#define AX_ATOMIC_MAX32(p, v) ax_host_atomic_max((int*)(p), (int)(v))
#endif

// ------------------------------------------------------------------ lookups
template <typename T>
AX_HD int ax_find_owner(const AxDrop* drop, int n_drops, T AxDrop::*base, int64_t v) {
    int lo = 0, hi = n_drops - 1;          // last drop with drop[i].base <= v
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if ((int64_t)(drop[mid].*base) <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// first index in [0,n) with a[i] >= v
AX_HD int64_t ax_lower_bound(const int32_t* a, int64_t n, int64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((int64_t)a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}
// first index in [0,n) with a[i] > v
AX_HD int64_t ax_upper_bound(const int32_t* a, int64_t n, int64_t v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if ((int64_t)a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------ stats
// reference AXCTDprocessor.py:55-56: np.mean (exact: integer sum < 2^53) and
// np.max(np.abs(int16)) where abs(-32768) wraps to -32768.
AX_HDN inline void ax_stats_item(const AxWave& w, int64_t slab) {
    int d = w.slab_drop[slab];
    const AxDrop& dr = w.drop[d];
    int64_t j = slab - dr.slab_base;
    int64_t a = j * AX_STAT_SLAB, b = a + AX_STAT_SLAB;
    if (b > dr.n) b = dr.n;
    const int16_t* x = w.pcm + dr.pcm_off;
    int64_t sum = 0;
    int mx = -32768;
    for (int64_t n = a; n < b; ++n) {
        int v = x[n];
        sum += v;
        int av = (v == -32768) ? -32768 : (v < 0 ? -v : v);
        if (av > mx) mx = av;
    }
    AX_ATOMIC_ADD64(&w.st[d].sum, sum);
    AX_ATOMIC_MAX32(&w.st[d].ampl, mx);
}

AX_HDN inline void ax_stats_fin(const AxWave& w, int64_t d) {
    AxState& st = w.st[d];
    st.dc = ax_div((double)st.sum, (double)w.drop[d].n);
    st.ampl_d = (double)st.ampl;
    st.inv_ampl = ax_div(1.0, st.ampl_d);
}

// ------------------------------------------------------------------ filter
// One biquad of the continuous pass (direct form II transposed as scipy's
// _sosfilt, but with FMAs: this pass only has to agree with the exact
// restart to ~1e-16 of peak; decisions inside `guard` of zero are flagged).
AX_HD double ax_biquad_fma(double x, const double* c, double& z0, double& z1) {
    double y = ax_fma(c[0], x, z0);
    z0 = ax_fma(c[1], x, ax_fma(-c[4], y, z1));
    z1 = ax_fma(c[2], x, -(c[5] * y));
    return y;
}
// scipy/signal/_sosfilt.pyx order, no contraction:
//   y = b0*x + z0 ; z0 = (b1*x - a1*y) + z1 ; z1 = b2*x - a2*y
AX_HD double ax_biquad_exact(double x, const double* c, double& z0, double& z1) {
    double y = ax_add(ax_mul(c[0], x), z0);
    z0 = ax_add(ax_sub(ax_mul(c[1], x), ax_mul(c[4], y)), z1);
    z1 = ax_sub(ax_mul(c[2], x), ax_mul(c[5], y));
    return y;
}

struct AxPending {
    double snap[AX_PEND][4];
    int64_t tgt[AX_PEND];
    int32_t idx[AX_PEND];
    int n;
};

// Continuous filter over one segment (with warm-up overlap), emitting every
// zero crossing i (sign(y[i]) != sign(y[i+1]), demodulate.py:77-79) together
// with |sum y[i+1..i+npcm] e^{j theta_f m}| for the mark and space tones
// (demodulate.py:99-102) via re-anchored prefix sums.
template <int NSEC>
AX_HDN inline void ax_filter_segment(const AxWave& w, int64_t seg) {
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    const int64_t L = w.seg_len;
    const int64_t j = seg - dr.seg_base;
    const int64_t seg_start = j * L;
    int64_t seg_end = seg_start + L;
    if (seg_end > dr.n) seg_end = dr.n;
    int64_t n_begin = seg_start - c.warm;
    if (n_begin < 0) n_begin = 0;
    int64_t n_stop = seg_end + c.npcm;       // last window of this segment ends at seg_end-1+npcm
    if (n_stop > dr.n) n_stop = dr.n;
    const int16_t* x = w.pcm + dr.pcm_off;
    const double kmul = st.inv_ampl, kadd = -(st.dc * st.inv_ampl);
    double cf[NSEC][6];
    double z[NSEC][2];
    for (int s = 0; s < NSEC; ++s) {
        for (int q = 0; q < 6; ++q) cf[s][q] = c.sos[s][q];
        z[s][0] = 0.0; z[s][1] = 0.0;
    }
    const int R = c.rebase, npcm = c.npcm;
    const double* tab = c.bit_cs;
    const double guard = w.guard;
    double C0 = 0, C1 = 0, C2 = 0, C3 = 0;
    AxPending pe;
    pe.n = 0;
    int m = 0, cnt = 0, unc = 0;
    bool prev_neg = false, have_prev = false;
    const int64_t slot = seg * (int64_t)w.seg_cap;
    for (int64_t n = n_begin; n < n_stop; ++n) {
        double u = ax_fma((double)x[n], kmul, kadd);
#pragma unroll
        for (int s = 0; s < NSEC; ++s) u = ax_biquad_fma(u, cf[s], z[s][0], z[s][1]);
        const bool neg = u < 0.0;
        if (have_prev && neg != prev_neg) {
            const int64_t i = n - 1;
            if (i >= seg_start && i < seg_end) {
                if (pe.n == AX_PEND) {       // too many crossings inside one bit window: give up on the oldest
                    if (cnt < w.seg_cap) { w.rec_idx[slot + cnt] = pe.idx[0]; w.rec_a1[slot + cnt] = ax_nan(); w.rec_a2[slot + cnt] = ax_nan(); }
                    ++cnt; ++unc;
                    for (int q = 1; q < AX_PEND; ++q) {
                        for (int r = 0; r < 4; ++r) pe.snap[q - 1][r] = pe.snap[q][r];
                        pe.tgt[q - 1] = pe.tgt[q]; pe.idx[q - 1] = pe.idx[q];
                    }
                    pe.n--;
                }
#pragma unroll
                for (int q = 0; q < AX_PEND; ++q) if (q == pe.n) {
                    pe.snap[q][0] = C0; pe.snap[q][1] = C1; pe.snap[q][2] = C2; pe.snap[q][3] = C3;
                    pe.tgt[q] = i + npcm; pe.idx[q] = (int32_t)i;
                }
                pe.n++;
            }
        }
        const double* t4 = tab + 4 * m;
        C0 = ax_fma(u, t4[0], C0); C1 = ax_fma(u, t4[1], C1);
        C2 = ax_fma(u, t4[2], C2); C3 = ax_fma(u, t4[3], C3);
        if (pe.n > 0 && pe.tgt[0] == n) {
            const double a1 = hypot(C0 - pe.snap[0][0], C1 - pe.snap[0][1]);
            const double a2 = hypot(C2 - pe.snap[0][2], C3 - pe.snap[0][3]);
            if (cnt < w.seg_cap) { w.rec_idx[slot + cnt] = pe.idx[0]; w.rec_a1[slot + cnt] = a1; w.rec_a2[slot + cnt] = a2; }
            ++cnt;
#pragma unroll
            for (int q = 1; q < AX_PEND; ++q) {
                for (int r = 0; r < 4; ++r) pe.snap[q - 1][r] = pe.snap[q][r];
                pe.tgt[q - 1] = pe.tgt[q]; pe.idx[q - 1] = pe.idx[q];
            }
            pe.n--;
        }
        if (n >= seg_start && n < seg_end && fabs(u) < guard) ++unc;
        if (++m == R) {
            // re-anchor the phase reference: snap' = -(C - snap) * e^{-j theta R}; C = 0
#pragma unroll
            for (int q = 0; q < AX_PEND; ++q) if (q < pe.n) {
                double pr = C0 - pe.snap[q][0], pi = C1 - pe.snap[q][1];
                pe.snap[q][0] = -(pr * c.rot[0][0] - pi * c.rot[0][1]);
                pe.snap[q][1] = -(pr * c.rot[0][1] + pi * c.rot[0][0]);
                pr = C2 - pe.snap[q][2]; pi = C3 - pe.snap[q][3];
                pe.snap[q][2] = -(pr * c.rot[1][0] - pi * c.rot[1][1]);
                pe.snap[q][3] = -(pr * c.rot[1][1] + pi * c.rot[1][0]);
            }
            C0 = C1 = C2 = C3 = 0.0;
            m = 0;
        }
        prev_neg = neg; have_prev = true;
    }
    // windows that run past the end of the recording can never be demodulated
    for (int q = 0; q < pe.n; ++q) {
        if (cnt < w.seg_cap) { w.rec_idx[slot + cnt] = pe.idx[q]; w.rec_a1[slot + cnt] = ax_nan(); w.rec_a2[slot + cnt] = ax_nan(); }
        ++cnt;
    }
    if (cnt > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; cnt = w.seg_cap; }
    w.seg_cnt[seg] = cnt;
    if (unc) AX_ATOMIC_ADD32(&st.n_uncertain, unc);
}

AX_HDN inline void ax_filter_item(const AxWave& w, int64_t seg) {
    const int nsec = w.cfg[w.drop[w.seg_drop[seg]].cfg].nsec;
    if (nsec == 3) ax_filter_segment<3>(w, seg);
    else if (nsec == 6) ax_filter_segment<6>(w, seg);
    else if (nsec == 1) ax_filter_segment<1>(w, seg);
    else if (nsec == 2) ax_filter_segment<2>(w, seg);
    else if (nsec == 4) ax_filter_segment<4>(w, seg);
    else ax_filter_segment<5>(w, seg);
}

// exclusive scan of the per-segment crossing counts of one drop
AX_HDN inline void ax_scan_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    int64_t off = 0;
    for (int s = 0; s < dr.nseg; ++s) { w.seg_off[dr.seg_base + s] = off; off += w.seg_cnt[dr.seg_base + s]; }
    w.st[d].zc_count = off;
    if (off > dr.zc_cap) { w.flags[AX_FLAG_CAP] = 1; w.st[d].zc_count = 0; ax_raise(w.st[d], AXCTD_DROP_CAPACITY, -1); }
}

AX_HDN inline void ax_compact_item(const AxWave& w, int64_t seg) {
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    if (w.st[d].zc_count == 0) return;
    const int64_t src = seg * (int64_t)w.seg_cap, dst = dr.zc_base + w.seg_off[seg];
    const int cnt = w.seg_cnt[seg];
    for (int q = 0; q < cnt; ++q) {
        w.zc_idx[dst + q] = w.rec_idx[src + q];
        w.zc_a1[dst + q] = w.rec_a1[src + q];
        w.zc_a2[dst + q] = w.rec_a2[src + q];
    }
}

// ------------------------------------------------------------------ walk
// demodulate.py:91-92: among the next four crossings take the one closest to
// one bit period later (first minimum).  |z - (z0 + fs/bitrate)| is compared
// as the integer |(z - z0)*2*bitrate - 2*fs| (exact whenever fs/bitrate is a
// dyadic rational, which holds for every rate divisible by 25).
AX_HD int64_t ax_next(const int32_t* zi, int64_t pos, int64_t fs2, int64_t br2) {
    const int64_t c0 = zi[pos];
    int64_t best = 0; int bj = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int64_t dd = ((int64_t)zi[pos + 1 + j] - c0) * br2 - fs2;
        if (dd < 0) dd = -dd;
        if (j == 0 || dd < best) { best = dd; bj = j; }
    }
    return pos + 1 + bj;
}

// tile_tab[(tile*4 + o)]: walk entering tile at offset o leaves it at offset
// (tab & 3) of the next tile after (tab >> 2) steps; 0xFFFF = not available.
AX_HDN inline void ax_tiles_item(const AxWave& w, int64_t item) {
    const int64_t tg = item >> 2;
    const int o = (int)(item & 3);
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::tile_base, tg);
    const AxDrop& dr = w.drop[d];
    const int64_t t = tg - dr.tile_base;
    if (t >= dr.tile_cap) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int64_t M = w.st[d].zc_count;
    const int32_t* zi = w.zc_idx + dr.zc_base;
    int64_t pos = t * AX_TILE + o;
    const int64_t limit = (t + 1) * AX_TILE;
    int cnt = 0;
    bool ok = pos < M;
    while (ok && pos < limit) {
        if (pos + 4 >= M) { ok = false; break; }
        pos = ax_next(zi, pos, c.fs2, 2 * (int64_t)c.bitrate);
        ++cnt;
    }
    w.tile_tab[tg * 4 + o] = ok ? (uint16_t)((pos - limit) | (cnt << 2)) : (uint16_t)0xFFFF;
}

// Follow the walk from `pos` while pos <= qstop-5 (demodulate.py:90:
// `while c < len(zerocrossings)-5`); returns the final ordinal.
AX_HD int64_t ax_walk_end(const int32_t* zi, const uint16_t* tab, int64_t pos, int64_t qstop,
                          int64_t fs2, int64_t br2, int64_t* steps) {
    int64_t st = 0;
    while (pos <= qstop - 5) {
        const int64_t t = pos / AX_TILE;
        const int o = (int)(pos - t * AX_TILE);
        if (o < 4 && (t + 1) * AX_TILE - 1 <= qstop - 5) {
            const uint16_t e = tab[t * 4 + o];
            if (e != 0xFFFF) { st += e >> 2; pos = (t + 1) * AX_TILE + (e & 3); continue; }
        }
        pos = ax_next(zi, pos, fs2, br2);
        ++st;
    }
    *steps = st;
    return pos;
}

// ------------------------------------------------------------------ chunk chain
// Predict the data-dependent chunk starts (AXCTDprocessor.py:327-329 with
// demodulate.py:104) from the continuous pass alone.
AX_HDN inline void ax_chain_item(const AxWave& w, int64_t d) {
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || st.chain_end) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int32_t* zi = w.zc_idx + dr.zc_base;
    const uint16_t* tab = w.tile_tab + (int64_t)dr.tile_base * 4;
    const int64_t M = st.zc_count;
    int k = st.chain_from;
    int64_t s;
    if (k == st.k0) s = ch[k].s;
    else s = ch[k - 1].true_last - 1 - c.pad;          // s + (last_edge - 1) - pad
    for (;; ++k) {
        if (dr.n - s < 4 * (int64_t)c.n_power) { st.n_chunks = k; break; }          // :295
        if (k >= dr.chunk_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; st.n_chunks = k; break; }
        int64_t e = s + c.chunk_len;                                               // :293
        if (e >= dr.n) e = dr.n - 1;                                               // :299-300
        ch[k].s = s; ch[k].e = e; ch[k].err = 0; ch[k].n_edges = 0; ch[k].spec_last = -1;
        const int64_t entry = ax_lower_bound(zi, M, s + c.pad);
        const int64_t q = ax_upper_bound(zi, M, e - 2) - 1;
        if (entry > q) { st.n_chunks = k + 1; break; }
        int64_t steps;
        const int64_t pos = ax_walk_end(zi, tab, entry, q, c.fs2, 2 * (int64_t)c.bitrate, &steps);
        ch[k].spec_last = zi[pos];
        const int64_t next_ind = zi[pos] - s - 1;                                  // demodulate.py:104
        if (next_ind <= c.pad) { st.n_chunks = k + 1; break; }                     // :330-331 handled by verify
        s = s + next_ind - c.pad;                                                  // :329
    }
}

// ------------------------------------------------------------------ exact head
// Recompute the first `head` samples of a chunk from zero filter state with
// scipy's exact operation order, find its crossings and bit edges, then join
// the precomputed continuous crossings.
AX_HDN inline void ax_head_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.status != 0 || st.sm_status < 1 || k < st.chain_from || k >= st.n_chunks || k >= dr.chunk_cap) return;
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk& ch = w.chunk[cg];
    const int64_t s = ch.s, e = ch.e, len = e - s;
    const int16_t* x = w.pcm + dr.pcm_off + s;
    const int64_t H = (w.force_exact || c.head > len) ? len : c.head;
    int64_t ny = H + c.npcm + 2;
    if (ny > len) ny = len;
    double* yb = w.ybuf + cg * (int64_t)w.ybuf_len_max;
    int32_t* hz = w.head_idx + cg * (int64_t)w.head_zc_cap_max;
    double* ha1 = w.head_a1 + cg * (int64_t)w.head_zc_cap_max;
    double* ha2 = w.head_a2 + cg * (int64_t)w.head_zc_cap_max;
    ch.err = 0; ch.n_edges = 0; ch.n_head_edges = 0; ch.g_first = -1; ch.true_last = -1; ch.q_last = -1; ch.first_edge = -1;
    if (ny > w.ybuf_len_max) { ch.err = AXCTD_DROP_CAPACITY; return; }
    {   // demodulate.py:74 on AXCTDprocessor.py:57 samples
        double z[AX_MAXSEC][2];
        for (int q = 0; q < AX_MAXSEC; ++q) { z[q][0] = 0.0; z[q][1] = 0.0; }
        for (int64_t n = 0; n < ny; ++n) {
            double u = ax_div(ax_sub((double)x[n], st.dc), st.ampl_d);
            for (int q = 0; q < c.nsec; ++q) u = ax_biquad_exact(u, c.sos[q], z[q][0], z[q][1]);
            yb[n] = u;
        }
    }
    int nh = 0;
    bool overflow = false;
    for (int64_t i = c.pad; i <= H - 2; ++i) {                 // demodulate.py:77-82
        if ((yb[i] < 0.0) != (yb[i + 1] < 0.0)) {
            if (nh < w.head_zc_cap_max) hz[nh++] = (int32_t)i; else overflow = true;
        }
    }
    if (overflow) { ch.err = AXCTD_DROP_CAPACITY; return; }
    const int32_t* zi = w.zc_idx + dr.zc_base;
    const uint16_t* tab = w.tile_tab + (int64_t)dr.tile_base * 4;
    const int64_t M = st.zc_count;
    int64_t g0 = 0, q = -1, nc = 0;
    if (H < len) {
        g0 = ax_lower_bound(zi, M, s + H - 1);
        q = ax_upper_bound(zi, M, e - 2) - 1;
        nc = q - g0 + 1;
        if (nc < 0) nc = 0;
    }
    ch.q_last = q;
    const int64_t total = nh + nc;
    if (total == 0) { ch.err = AXCTD_DROP_NO_CROSSING; return; }      // demodulate.py:85
    const int64_t br2 = 2 * (int64_t)c.bitrate;
    ch.first_edge = (nh > 0) ? (s + hz[0]) : (int64_t)zi[g0];
    int64_t cpos = 0;
    int nhe = 0;
    bool done = false;
    int64_t last = -1;
    while (cpos < nh) {
        hz[nhe++] = hz[cpos];                                  // bit edge inside the head
        if (!(cpos < total - 5)) { last = s + hz[nhe - 1]; done = true; break; }
        const int64_t c0 = s + hz[cpos];
        int64_t best = 0; int bj = 0;
        for (int jj = 0; jj < 4; ++jj) {                       // demodulate.py:91-92
            const int64_t mj = cpos + 1 + jj;
            const int64_t v = (mj < nh) ? (s + hz[mj]) : (int64_t)zi[g0 + (mj - nh)];
            int64_t dd = (v - c0) * br2 - c.fs2;
            if (dd < 0) dd = -dd;
            if (jj == 0 || dd < best) { best = dd; bj = jj; }
        }
        cpos += 1 + bj;
    }
    int64_t nedges = nhe;
    if (!done) {
        const int64_t pos = g0 + (cpos - nh);
        ch.g_first = pos;
        int64_t steps;
        const int64_t endpos = ax_walk_end(zi, tab, pos, q, c.fs2, br2, &steps);
        nedges += steps + 1;
        last = zi[endpos];
    }
    ch.n_head_edges = nhe;
    ch.n_edges = (int32_t)nedges;
    ch.true_last = last;
    // demodulate.py:99-102 for the head edges (all but the chunk's final edge)
    for (int t = 0; t < nhe; ++t) {
        const int64_t e0 = hz[t];
        const bool is_last = (t == nedges - 1);
        if (e0 + c.inset + c.npcm > len) {
            ha1[t] = ax_nan(); ha2[t] = ax_nan();
            if (!is_last) ch.err = AXCTD_DROP_SHORT_WINDOW;
            continue;
        }
        double sr1 = 0, si1 = 0, sr2 = 0, si2 = 0;
        for (int mI = 0; mI < c.npcm; ++mI) {
            const double y = yb[e0 + c.inset + mI];
            const double* t4 = c.bit_cs + 4 * mI;
            sr1 = ax_fma(y, t4[0], sr1); si1 = ax_fma(y, t4[1], si1);
            sr2 = ax_fma(y, t4[2], sr2); si2 = ax_fma(y, t4[3], si2);
        }
        ha1[t] = hypot(sr1, si1);
        ha2[t] = hypot(sr2, si2);
    }
}

// Compare the prediction with the exact result; repair the chain on mismatch.
AX_HDN inline void ax_verify_item(const AxWave& w, int64_t d) {
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || st.chain_end) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    st.chain_dirty = 0;
    for (int k = st.chain_from; k < st.n_chunks; ++k) {
        if (ch[k].err) { ax_raise(st, ch[k].err, k); st.n_chunks = k + 1; st.chain_from = k + 1; st.chain_end = 1; return; }
        const int64_t next_ind = ch[k].true_last - ch[k].s - 1;
        if (next_ind <= c.pad) {
            // AXCTDprocessor.py:331: the start index becomes a float; the loop only survives
            // if the end-of-file test (:295) stops it first.
            const double sf = (double)ch[k].s + c.fs / (double)c.bitrate;
            if (!((double)dr.n - sf < 4.0 * c.n_power)) ax_raise(st, AXCTD_DROP_FLOAT_INDEX, k + 1);
            st.n_chunks = k + 1; st.chain_from = k + 1; st.chain_end = 1;
            return;
        }
        if (ch[k].true_last != ch[k].spec_last) {
            st.n_fixups++;
            st.chain_from = k + 1;
            st.chain_dirty = 1;
            w.flags[AX_FLAG_DIRTY] = 1;
            return;
        }
    }
    st.chain_from = st.n_chunks;
    st.chain_end = 1;
}
