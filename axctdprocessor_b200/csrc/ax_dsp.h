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
//   ax_headfilt_item / ax_headwalk_item   zero-state restart per chunk  demodulate.py:74 (sosfilt zero state)
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

// ------------------------------------------------------------------ guard band
// A filter output closer to zero than `guard` cannot be signed reliably by the fast passes (their operation order
// differs from scipy's by ~1e-14 of full scale).  Such samples are listed here and settled after the chunk chain is
// final (ax_unc_resolve_item): kp1 = 0 for the continuous pass, chunk + 1 for the zero-state head of that iteration.
AX_HD void ax_unc_push(const AxWave& w, int d, int64_t n, bool neg, int kp1, int64_t chunk_s) {
    const int slot = (int)AX_ATOMIC_ADD32(&w.st[d].n_unc_listed, 1);
    if (slot >= 0 && slot < AX_UNC_CAP) {
        int64_t* e = w.unc_list + ((int64_t)d * AX_UNC_CAP + slot) * 2;
        e[0] = (n & 0xffffffffll) | (neg ? (1ll << 32) : 0ll) | ((int64_t)kp1 << 33);
        e[1] = chunk_s;
    }
}

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
    if (b > dr.n_raw) b = dr.n_raw;
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
    // min / max decide np.max(np.abs(x)) unless a sample equals -32768 (then st.ampl holds the exact rescan)
    if (st.vmin != 0x7fffffff && st.vmin > -32768) st.ampl = st.vmax > -st.vmin ? st.vmax : -st.vmin;
    st.dc = ax_div((double)st.sum, (double)w.drop[d].n_raw);
    st.ampl_d = (double)st.ampl;
    st.inv_ampl = ax_div(1.0, st.ampl_d);
}

// ------------------------------------------------------------------ filter
// scipy/signal/_sosfilt.pyx order, no contraction (used by the exact heads):
//   y = b0*x + z0 ; z0 = (b1*x - a1*y) + z1 ; z1 = b2*x - a2*y
AX_HD double ax_biquad_exact(double x, const double* c, double& z0, double& z1) {
    double y = ax_add(ax_mul(c[0], x), z0);
    z0 = ax_add(ax_sub(ax_mul(c[1], x), ax_mul(c[4], y)), z1);
    z1 = ax_sub(ax_mul(c[2], x), ax_mul(c[5], y));
    return y;
}

// every section b2 == b0, b1 == +-2*b0, and b0 == 1 except in section 0
// (scipy.signal.butter(..., output='sos') for low- and band-pass)
AX_HD bool ax_sos_is_butter(const AxCfg& c) {
    for (int s = 0; s < c.nsec; ++s) {
        const double b0 = c.sos[s][0], b1 = c.sos[s][1], b2 = c.sos[s][2];
        if (b2 != b0 || (b1 != 2.0 * b0 && b1 != -2.0 * b0)) return false;
        if (s > 0 && b0 != 1.0) return false;
    }
    return true;
}

// ------------------------------------------------------------------ bit windows
// demodulate.py:99-102: |sum_m y[i+1+m] e^{j theta_f m}| for the mark and space tones over the
// npcm samples after crossing i.  Only the magnitude is used, so the phase reference is free:
// the window is evaluated over the nt = 4 * ax_win_quads(npcm) samples of the aligned quads that
// cover it, the samples outside the window masked to zero, with the phase reference in the MIDDLE
// of those nt taps: taps k and nt-1-k then carry conjugate phasors e^{-+j phi_k}, phi_k =
// theta_f (nt/2 - 1/2 - k), so that
//   Re S = sum_{k < nt/2} (y[k] + y[nt-1-k]) cos phi_k,   Im S = sum_{k < nt/2} (y[nt-1-k] - y[k]) sin phi_k:
// two additions and four FMAs per PAIR of samples for both tones (the sum and the difference serve
// both) instead of eight FMAs, and half the phasor loads.  tab[k], k < nt/2, holds (cos phi_k,
// sin phi_k) of the mark and of the space tone (axctd_config_create).  The samples are the float
// roundings of the double-precision filter output and the sums run in fp32 (eight chains); a
// decision that this precision cannot make is re-made from ax_gwin_* in double.
AX_HD int ax_win_quads(int npcm) { return (npcm + 6) >> 2; }     // quads covering offset o <= 3 plus npcm taps

// yv[k], k < 4*ax_win_quads(npcm): samples of the aligned quads; o = (i + 1) & 3
AX_HD void ax_window32(const float* yv, int o, int npcm, const AxF4* tab, float* a1, float* a2) {
    float r1a = 0.f, i1a = 0.f, r2a = 0.f, i2a = 0.f, r1b = 0.f, i1b = 0.f, r2b = 0.f, i2b = 0.f;
    const int nt = 4 * ax_win_quads(npcm), nh = nt >> 1;        // nh is even
    // tap k lies inside the window iff o <= k < o + npcm (o <= 3: only the first three and the last few taps can fall outside)
#define AX_WIN_TAP(k) ((((k) >= 3 || (k) >= o) && ((k) < npcm || (k) < o + npcm)) ? yv[(k)] : 0.f)
#pragma unroll
    for (int k = 0; k < AX_WIN_TAPS / 2; k += 2) {
        if (k < nh) {
            const float la = AX_WIN_TAP(k), ua = AX_WIN_TAP(nt - 1 - k);
            const float lb = AX_WIN_TAP(k + 1), ub = AX_WIN_TAP(nt - 2 - k);
            const float sa = la + ua, da = ua - la, sb = lb + ub, db = ub - lb;
            const AxF4 ta = tab[k], tb = tab[k + 1];                   // 16-byte broadcast loads (shared memory in k_demod_fused)
            r1a = fmaf(sa, ta.x, r1a); i1a = fmaf(da, ta.y, i1a);
            r2a = fmaf(sa, ta.z, r2a); i2a = fmaf(da, ta.w, i2a);
            r1b = fmaf(sb, tb.x, r1b); i1b = fmaf(db, tb.y, i1b);
            r2b = fmaf(sb, tb.z, r2b); i2b = fmaf(db, tb.w, i2b);
        }
    }
#undef AX_WIN_TAP
    const float r1 = r1a + r1b, i1 = i1a + i1b, r2 = r2a + r2b, i2 = i2a + i2b;
    *a1 = sqrtf(fmaf(r1, r1, i1 * i1));
    *a2 = sqrtf(fmaf(r2, r2, i2 * i2));
}

// The same two magnitudes in double precision, straight from the int16 samples: with h the
// impulse response of the SOS cascade and u = (x - dc) / ampl,
//   S_f(i) = sum_m e^{j theta_f m} y[i+1+m] = sum_{d >= 0} u[i + npcm - d] * G_f[d],
//   G_f[d] = sum_m e^{j theta_f m} h[d - (npcm - 1 - m)]            (host, long double)
// truncated at the sample q0 where the filter state was zero (the start of the recording or, for
// the reference's per-chunk restart, the chunk start).  Partial sums over d = lane, lane+nl, ...
AX_HD void ax_gwin_partial(const AxSrc& x, int64_t i, int64_t q0, const AxCfg& c, int lane, int nl, double* acc) {
    const int64_t n_end = i + c.npcm;
    int64_t D = n_end - q0;
    if (D > (int64_t)c.g_len - 1) D = (int64_t)c.g_len - 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int64_t d = lane; d <= D; d += nl) {
        const double xv = ax_get(x, n_end - d);
        const double* g = c.gtab + 4 * d;
        s0 = ax_fma(xv, g[0], s0); s1 = ax_fma(xv, g[1], s1); s2 = ax_fma(xv, g[2], s2); s3 = ax_fma(xv, g[3], s3);
    }
    acc[0] = s0; acc[1] = s1; acc[2] = s2; acc[3] = s3;
}
AX_HD void ax_gwin_finish(const double* acc, int64_t i, int64_t q0, const AxCfg& c, const AxState& st, double* a1, double* a2) {
    int64_t D = i + c.npcm - q0;
    if (D > (int64_t)c.g_len - 1) D = (int64_t)c.g_len - 1;
    const double* gc = c.gcum + 4 * D;
    double S[4];
    for (int q = 0; q < 4; ++q) S[q] = (acc[q] - st.dc * gc[q]) * st.inv_ampl;
    *a1 = hypot(S[0], S[1]);
    *a2 = hypot(S[2], S[3]);
}

// Per-thread state of the continuous filter pass over one segment (generic form; the sm_100a
// production kernel is k_demod_fused in ax_kernels.cuh, which computes the same values): the SOS
// cascade (direct form II transposed as scipy's _sosfilt, with FMAs: this pass only has to agree
// with the exact zero-state restart to ~1e-16 of peak, and samples inside `guard` of zero are
// flagged), the zero-crossing detector (demodulate.py:77-79) and the float bit windows above.
// BUTTER: 4 double-precision operations per section (5 in section 0, which also absorbs the
// (x - mean)/max|x| normalisation of AXCTDprocessor.py:57).
template <int NSEC, bool BUTTER>
struct AxFilt {
    double z0[NSEC], z1[NSEC], a1[NSEC], a2[NSEC];
    double sg[NSEC];                 // BUTTER: b1/b0 = +-2
    double b0[NSEC], b1[NSEC], b2[NSEC];   // general form
    double k0, k1;
    float yring[128];                // float filter output, index n & 127
    int32_t tgt[AX_PEND], idx[AX_PEND];
    int32_t np, head, next_tgt, cnt, unc, npcm, cap;   // pending windows: ring in local memory
    int32_t seg_start, seg_end;
    bool prev_neg, have_prev;
    double guard;
    const AxWinTab* tab;
    int32_t* rec_idx; float* rec_a1; float* rec_a2;
    const AxWave* uw; int32_t ud, ukp1; int64_t us;       // where guard-band samples are listed (ax_unc_push)

    AX_HD void init(const AxCfg& c, const AxState& st, int32_t s0, int32_t s1, double guard_,
                    int32_t* ri, float* r1, float* r2, int32_t cap_, const AxWave* uw_, int ud_, int ukp1_, int64_t us_) {
        uw = uw_; ud = ud_; ukp1 = ukp1_; us = us_;
        const double kmul = st.inv_ampl, kadd = -(st.dc * st.inv_ampl);
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            z0[s] = 0.0; z1[s] = 0.0;
            a1[s] = c.sos[s][4]; a2[s] = c.sos[s][5];
            b0[s] = c.sos[s][0]; b1[s] = c.sos[s][1]; b2[s] = c.sos[s][2];
            sg[s] = (c.sos[s][1] < 0.0) ? -2.0 : 2.0;
        }
        if (BUTTER) { k0 = c.sos[0][0] * kmul; k1 = c.sos[0][0] * kadd; } else { k0 = kmul; k1 = kadd; }
        for (int q = 0; q < 128; ++q) yring[q] = 0.f;
        np = 0; head = 0; next_tgt = 0x7fffffff; cnt = 0; unc = 0; npcm = c.npcm; cap = cap_;
        seg_start = s0; seg_end = s1; prev_neg = false; have_prev = false; guard = guard_;
        tab = &c.win_tab;
        rec_idx = ri; rec_a1 = r1; rec_a2 = r2;
    }

    AX_HD double filter(double xd) {
        if (BUTTER) {
            double t = ax_fma(xd, k0, k1);
#pragma unroll
            for (int s = 0; s < NSEC; ++s) {
                const double y = t + z0[s];
                z0[s] = ax_fma(-a1[s], y, ax_fma(sg[s], t, z1[s]));
                z1[s] = ax_fma(-a2[s], y, t);
                t = y;
            }
            return t;
        } else {
            double u = ax_fma(xd, k0, k1);
#pragma unroll
            for (int s = 0; s < NSEC; ++s) {
                const double y = ax_fma(b0[s], u, z0[s]);
                z0[s] = ax_fma(b1[s], u, ax_fma(-a1[s], y, z1[s]));
                z1[s] = ax_fma(b2[s], u, -(a2[s] * y));
                u = y;
            }
            return u;
        }
    }

    AX_HD void put(int32_t i, float v1, float v2) {
        if (cnt < cap) { rec_idx[cnt] = i; rec_a1[cnt] = v1; rec_a2[cnt] = v2; }
        ++cnt;
    }
    AX_HD void pop() {
        head = (head + 1) & (AX_PEND - 1);
        np--;
        next_tgt = np > 0 ? tgt[head] : 0x7fffffff;
    }

    // one sample: n = index within the drop, xd = (double) int16 sample
    AX_HD void step(int32_t n, double xd) {
        const double u = filter(xd);
        const bool neg = u < 0.0;
        if (have_prev && neg != prev_neg) {
            const int32_t i = n - 1;
            if (i >= seg_start && i < seg_end) {
                if (np == AX_PEND) { put(idx[head], (float)ax_nan(), (float)ax_nan()); ++unc; pop(); }   // cannot happen for a 1200 Hz band limit
                const int q = (head + np) & (AX_PEND - 1);
                tgt[q] = i + npcm; idx[q] = i;
                if (np == 0) next_tgt = i + npcm;
                np++;
            }
        }
        yring[n & 127] = (float)u;
        if (next_tgt == n) {
            const int32_t i = idx[head];
            const int32_t a = (i + 1) & ~3, o = (i + 1) & 3;
            float yv[AX_WIN_TAPS];
            const int nt = 4 * ax_win_quads(npcm);
            for (int k = 0; k < nt; ++k) yv[k] = yring[(a + k) & 127];
            float v1, v2;
            ax_window32(yv, o, npcm, tab->t, &v1, &v2);
            put(i, v1, v2);
            pop();
        }
        if (fabs(u) < guard && n >= seg_start && n < seg_end) { ++unc; ax_unc_push(*uw, ud, n, signbit(u) != 0, ukp1, us); }
        prev_neg = neg; have_prev = true;
    }

    // windows that run past the end of the recording can never be demodulated
    AX_HD void finish() {
        while (np > 0) { put(idx[head], (float)ax_nan(), (float)ax_nan()); pop(); }
    }
};

struct AxSegGeom { int64_t seg_start, seg_end, n_begin, n_stop; };
AX_HD AxSegGeom ax_seg_geom(const AxDrop& dr, const AxCfg& c, int64_t L, int64_t j) {
    AxSegGeom g;
    g.seg_start = j * L;
    g.seg_end = g.seg_start + L; if (g.seg_end > dr.n) g.seg_end = dr.n;
    g.n_begin = g.seg_start - c.warm; if (g.n_begin < 0) g.n_begin = 0;
    g.n_stop = g.seg_end + c.npcm; if (g.n_stop > dr.n) g.n_stop = dr.n;      // last window of the segment ends at seg_end-1+npcm
    return g;
}

template <int NSEC, bool BUTTER>
AX_HDN inline void ax_filter_segment(const AxWave& w, int64_t seg) {
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    const int64_t j = seg - dr.seg_base;
    if (j >= dr.nseg) { w.seg_cnt[seg] = 0; w.seg_unc[seg] = 0; return; }
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    if (w.streaming && j < st.seg_done) return;          // its records are final (an earlier run of the growing recording)
    const AxSegGeom g = ax_seg_geom(dr, c, w.seg_len, j);
    const AxSrc x = ax_src(w, dr);
    const int64_t slot = seg * (int64_t)w.seg_cap;
    AxFilt<NSEC, BUTTER> f;
    f.init(c, st, (int32_t)g.seg_start, (int32_t)g.seg_end, w.guard, w.rec_idx + slot, w.rec_a1 + slot, w.rec_a2 + slot, w.seg_cap,
           &w, d, 0, 0);
    for (int64_t n = g.n_begin; n < g.n_stop; ++n) f.step((int32_t)n, ax_get(x, n));
    f.finish();
    if (f.cnt > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; ax_raise(st, AXCTD_DROP_CAPACITY, -1); f.cnt = w.seg_cap; }   // crossings were dropped: fail the drop
    w.seg_cnt[seg] = f.cnt;
    w.seg_unc[seg] = f.unc;
}

AX_HDN inline void ax_filter_item(const AxWave& w, int64_t seg) {
    if (w.only_xf && w.drop[w.seg_drop[seg]].xf_off < 0) return;      // the fused kernel filtered the int16 drops
    const AxCfg& c = w.cfg[w.drop[w.seg_drop[seg]].cfg];
    const bool bt = ax_sos_is_butter(c);
    if (c.nsec == 3) { if (bt) ax_filter_segment<3, true>(w, seg); else ax_filter_segment<3, false>(w, seg); }
    else if (c.nsec == 6) { if (bt) ax_filter_segment<6, true>(w, seg); else ax_filter_segment<6, false>(w, seg); }
    else if (c.nsec == 1) ax_filter_segment<1, false>(w, seg);
    else if (c.nsec == 2) ax_filter_segment<2, false>(w, seg);
    else if (c.nsec == 4) ax_filter_segment<4, false>(w, seg);
    else ax_filter_segment<5, false>(w, seg);
}

// exclusive scan of the crossing counts, stage 1: inside blocks of 128 segments
AX_HDN inline void ax_scan_block_item(const AxWave& w, int64_t blk) {
    int64_t off = 0;
    const int64_t s0 = blk * 128;
    for (int q = 0; q < 128; ++q) { w.seg_off[s0 + q] = off; off += w.seg_cnt[s0 + q]; }
    w.blk_sum[blk] = off;
}

// exclusive scan of the per-segment crossing counts of one drop
AX_HDN inline void ax_scan_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    int64_t off = 0;                                   // stage 2: across the drop's blocks (seg_base is a multiple of 128)
    const int nblk = (dr.nseg + 127) / 128, b0 = dr.seg_base / 128;
    for (int q = 0; q < nblk; ++q) { const int64_t t = w.blk_sum[b0 + q]; w.blk_sum[b0 + q] = off; off += t; }
    w.st[d].zc_count = off;
    if (off > dr.zc_cap) { w.flags[AX_FLAG_CAP] = 1; w.st[d].zc_count = 0; ax_raise(w.st[d], AXCTD_DROP_CAPACITY, -1); }
}

AX_HDN inline void ax_compact_item(const AxWave& w, int64_t seg) {
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    if (w.st[d].zc_count == 0) return;
    const int64_t src = seg * (int64_t)w.seg_cap, dst = dr.zc_base + w.seg_off[seg] + w.blk_sum[seg / 128];
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

// zc_nx[c] = ax_next(c) - c (1..4), 0 where the walk cannot step (fewer than four crossings follow)
AX_HDN inline void ax_nx_item(const AxWave& w, int64_t zg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::zc_base, zg);
    const AxDrop& dr = w.drop[d];
    const int64_t pos = zg - dr.zc_base;
    const int64_t M = w.st[d].zc_count;
    if (pos >= M) return;
    const AxCfg& c = w.cfg[dr.cfg];
    w.zc_nx[zg] = (pos + 4 < M) ? (uint8_t)(ax_next(w.zc_idx + dr.zc_base, pos, c.fs2, 2 * (int64_t)c.bitrate) - pos) : (uint8_t)0;
}

// Per tile of AX_TILE crossings:
//   tile_mask[t][o]  bit i set <=> crossing t*AX_TILE+i is visited by the walk entering the tile at offset o (0..3)
//   tile_map[t]      byte o = offset (0..3) at which that walk enters the next tile, 0xFF if it stops inside
AX_HDN inline void ax_tiles_item(const AxWave& w, int64_t tg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::tile_base, tg);
    const AxDrop& dr = w.drop[d];
    const int64_t t = tg - dr.tile_base;
    if (t >= dr.tile_cap) return;
    const int64_t M = w.st[d].zc_count;
    const int64_t first = t * AX_TILE;
    if (first >= M) return;
    const uint8_t* nx = w.zc_nx + dr.zc_base;
    const int64_t limit = first + AX_TILE;
    uint32_t map = 0;
    for (int o = 0; o < 4; ++o) {
        uint64_t mask = 0;
        uint32_t ex = 0xFF;
        int64_t c = first + o;
        while (c < M) {
            if (c >= limit) { ex = (uint32_t)(c - limit); break; }
            mask |= 1ull << (c - first);
            if (!nx[c]) break;
            c += nx[c];
        }
        w.tile_mask[tg * 4 + o] = mask;
        map |= ex << (8 * o);
    }
    w.tile_map[tg] = map;
}

AX_HD int ax_popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
AX_HD int ax_ctz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}

// ---- the canonical walk -------------------------------------------------------------------------
// The greedy walk of every demodulated chunk starts at (or right beside) the crossing where the
// previous chunk's walk ended, so apart from the exact heads all chunks of a drop travel along ONE
// walk: the one that starts at the first usable crossing of the first demodulated chunk.  Its visited
// set is stored as one 64-bit mask per tile (cmask) with the running count of visited crossings
// before each tile (crank).  A walk that starts elsewhere is stepped explicitly until it lands on
// a canonical crossing; from there "where does it stop" and "how many steps" are two lookups.
//
// First tile of one drop: returns the state (offset into the next tile, 4 = stopped) and fills
// cmask[t0] for the walk that starts at ordinal `entry`.
AX_HD int ax_canon_first_tile(const uint8_t* nx, int64_t M, int64_t entry, uint64_t* mask_out) {
    const int64_t first = (entry / AX_TILE) * AX_TILE, limit = first + AX_TILE;
    uint64_t mask = 0;
    int state = 4;
    int64_t c = entry;
    while (c < M) {
        if (c >= limit) { state = (int)(c - limit); break; }
        mask |= 1ull << (c - first);
        if (!nx[c]) break;
        c += nx[c];
    }
    *mask_out = mask;
    return state;
}
AX_HD int ax_map_apply(uint32_t map, int state) {
    if (state >= 4) return 4;
    const uint32_t e = (map >> (8 * state)) & 0xFFu;
    return e == 0xFFu ? 4 : (int)e;
}

// generic (sequential) form of the canonical-walk tables of one drop; the CUDA build uses k_canon_block
AX_HDN inline void ax_canon_item(const AxWave& w, int64_t d) {
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int64_t M = st.zc_count;
    const int64_t ntile = (M + AX_TILE - 1) / AX_TILE;
    uint64_t* cmask = w.cmask + dr.tile_base;
    int32_t* crank = w.crank + dr.tile_base;
    const int64_t entry = ax_lower_bound(w.zc_idx + dr.zc_base, M, w.chunk[dr.chunk_base + st.k0].s + c.pad);
    const int64_t t0 = entry / AX_TILE;
    for (int64_t t = 0; t < ntile && t < t0; ++t) { cmask[t] = 0; crank[t] = 0; }
    if (entry >= M) return;
    uint64_t m0;
    int state = ax_canon_first_tile(w.zc_nx + dr.zc_base, M, entry, &m0);
    cmask[t0] = m0; crank[t0] = 0;
    int32_t rank = ax_popc64(m0);
    for (int64_t t = t0 + 1; t < ntile; ++t) {
        const uint64_t m = state < 4 ? w.tile_mask[(dr.tile_base + t) * 4 + state] : 0ull;
        cmask[t] = m; crank[t] = rank;
        rank += ax_popc64(m);
        state = ax_map_apply(w.tile_map[dr.tile_base + t], state);
    }
}

// number of canonical crossings with ordinal < pos
AX_HD int64_t ax_canon_rank(const uint64_t* cmask, const int32_t* crank, int64_t pos) {
    const int64_t t = pos / AX_TILE;
    return (int64_t)crank[t] + ax_popc64(cmask[t] & ((1ull << (pos - t * AX_TILE)) - 1ull));
}

// Follow the walk from `pos` while pos <= qstop-5 (demodulate.py:90:
// `while c < len(zerocrossings)-5`); returns the final ordinal and the number of steps.
// *merge (optional): ordinal at which the walk joined the canonical walk, -1 if it ended before.
AX_HD int64_t ax_walk_end(const uint8_t* nx, const uint64_t* cmask, const int32_t* crank, int64_t pos, int64_t qstop,
                          int64_t* steps, int64_t* merge = nullptr) {
    int64_t st = 0;
    if (merge) *merge = -1;
    const int64_t X = qstop - 4;                       // the walk stops at the first visited ordinal >= X
    while (pos < X) {
        const int64_t t = pos / AX_TILE;
        if ((cmask[t] >> (pos - t * AX_TILE)) & 1ull) {  // on the canonical walk: jump
            if (merge) *merge = pos;
            int64_t tx = X / AX_TILE;
            uint64_t m = cmask[tx] & ~((1ull << (X - tx * AX_TILE)) - 1ull);
            while (!m) m = cmask[++tx];                 // (the canonical walk passes every ordinal range [X, X+3] with X <= M-5)
            const int64_t endp = tx * AX_TILE + ax_ctz64(m);
            st += ax_canon_rank(cmask, crank, endp) - ax_canon_rank(cmask, crank, pos);
            pos = endp;
            break;
        }
        pos += nx[pos];                                // (nx > 0 is guaranteed while pos <= qstop-5)
        ++st;
    }
    *steps = st;
    return pos;
}

// first index in [0,n) with a[i] > v, searching outward from a guess
AX_HD int64_t ax_upper_bound_from(const int32_t* a, int64_t n, int64_t v, int64_t guess) {
    if (n <= 0) return 0;
    if (guess < 0) guess = 0;
    if (guess > n - 1) guess = n - 1;
    int64_t lo, hi;                                    // invariant: a[lo-1] <= v (or lo == 0), a[hi] > v (or hi == n)
    if ((int64_t)a[guess] <= v) {
        int64_t step = 1; lo = guess + 1; hi = lo;
        while (hi < n && (int64_t)a[hi] <= v) { lo = hi + 1; hi += step; step <<= 1; }
        if (hi > n) hi = n;
    } else {
        int64_t step = 1; hi = guess; lo = guess;
        while (lo > 0 && (int64_t)a[lo - 1] > v) { hi = lo - 1; lo -= step; step <<= 1; if (lo < 0) lo = 0; }
    }
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((int64_t)a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
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
    const uint8_t* nx = w.zc_nx + dr.zc_base;
    const uint64_t* cmask = w.cmask + dr.tile_base;
    const int32_t* crank = w.crank + dr.tile_base;
    const int64_t M = st.zc_count;
    int k = st.chain_from;
    int64_t s;
    if (k == st.k0) s = ch[k].s;
    else s = ch[k - 1].true_last - 1 - c.pad;          // s + (last_edge - 1) - pad
    int64_t entry = -1, span = (int64_t)c.chunk_len / 37;
    for (;; ++k) {
        if (w.streaming == 1) { if (s + c.chunk_len >= dr.n) { st.n_chunks = k; break; } }   // not complete yet: a later run takes it
        else if (dr.n - s < 4 * (int64_t)c.n_power) { st.n_chunks = k; break; }     // :295
        if (k >= dr.chunk_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; st.n_chunks = k; break; }
        int64_t e = s + c.chunk_len;                                               // :293
        if (e >= dr.n) e = dr.n - 1;                                               // :299-300
        ch[k].s = s; ch[k].e = e; ch[k].err = 0; ch[k].n_edges = 0; ch[k].spec_last = -1;
        if (entry < 0) entry = ax_lower_bound(zi, M, s + c.pad);
        const int64_t q = ax_upper_bound_from(zi, M, e - 2, entry + span) - 1;
        if (entry > q) { st.n_chunks = k + 1; break; }
        span = q - entry;
        int64_t steps;
        const int64_t pos = ax_walk_end(nx, cmask, crank, entry, q, &steps);
        ch[k].spec_last = zi[pos];
        const int64_t next_ind = zi[pos] - s - 1;                                  // demodulate.py:104
        if (next_ind <= c.pad) { st.n_chunks = k + 1; break; }                     // :330-331 handled by verify
        s = s + next_ind - c.pad;                                                  // :329
        // next entry: first crossing with index >= s + pad = zi[pos] - 1
        entry = (pos > 0 && (int64_t)zi[pos - 1] >= (int64_t)zi[pos] - 1) ? pos - 1 : pos;
    }
}

// ------------------------------------------------------------------ chunk heads
// The reference restarts its filter from zero state at every chunk start (demodulate.py:74 on the chunk
// slice).  Past the first `head` samples that restart is indistinguishable from the continuous pass; the
// head itself is filtered again from zero state, with the same arithmetic as the continuous pass
// (AxFilt / k_demod_fused in head mode), giving the crossings pad <= i <= H-2 (demodulate.py:77-82) and
// their windows.
struct AxHeadGeom { int64_t s, len, H, ny; bool active; };
AX_HD AxHeadGeom ax_head_geom(const AxWave& w, const AxDrop& dr, const AxState& st, const AxCfg& c, const AxChunk& ch, int k) {
    AxHeadGeom g;
    g.active = !(st.status != 0 || st.sm_status < 1 || k < st.chain_from || k >= st.n_chunks || k >= dr.chunk_cap);
    g.s = ch.s; g.len = ch.e - ch.s;
    g.H = (w.force_exact || c.head > g.len) ? g.len : c.head;
    g.ny = g.H + c.npcm + 2;
    if (g.ny > g.len) g.ny = g.len;
    return g;
}

// generic form of the head filter (the CUDA build runs k_demod_fused<.., HEAD> for the rate classes it
// is instantiated for; `only_rest`: skip those)
template <int NSEC, bool BUTTER>
AX_HDN inline void ax_headfilt_run(const AxWave& w, int64_t cg, int d, const AxDrop& dr, const AxCfg& c, AxState& st, const AxHeadGeom& g) {
    int32_t* hz = w.head_idx + cg * (int64_t)w.head_zc_cap_max;
    float* ha1 = w.head_a1 + cg * (int64_t)w.head_zc_cap_max;
    float* ha2 = w.head_a2 + cg * (int64_t)w.head_zc_cap_max;
    const AxSrc x = ax_src(w, dr);
    AxFilt<NSEC, BUTTER> f;
    f.init(c, st, (int32_t)(g.s + c.pad), (int32_t)(g.s + g.H - 1), w.guard, hz, ha1, ha2, w.head_zc_cap_max,
           &w, d, (int)(cg - dr.chunk_base) + 1, g.s);
    for (int64_t n = g.s; n < g.s + g.ny; ++n) f.step((int32_t)n, ax_get(x, n));
    f.finish();
    const int cnt = f.cnt > w.head_zc_cap_max ? -1 : f.cnt;
    for (int q = 0; q < cnt; ++q) hz[q] -= (int32_t)g.s;               // chunk-relative
    w.head_cnt[cg] = cnt;
    w.head_unc[cg] = f.unc;
}
AX_HDN inline void ax_headfilt_item(const AxWave& w, int64_t cg, int only_rest) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (k >= dr.chunk_cap) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const AxHeadGeom g = ax_head_geom(w, dr, st, c, w.chunk[cg], k);
    if (!g.active) return;
    const bool bt = ax_sos_is_butter(c);
    if (only_rest && bt && (c.nsec == 3 || c.nsec == 6) && (c.npcm == 39 || c.npcm == 43) && c.inset == 1 && dr.xf_off < 0) return;
    if (c.nsec == 3) { if (bt) ax_headfilt_run<3, true>(w, cg, d, dr, c, st, g); else ax_headfilt_run<3, false>(w, cg, d, dr, c, st, g); }
    else if (c.nsec == 6) { if (bt) ax_headfilt_run<6, true>(w, cg, d, dr, c, st, g); else ax_headfilt_run<6, false>(w, cg, d, dr, c, st, g); }
    else if (c.nsec == 1) ax_headfilt_run<1, false>(w, cg, d, dr, c, st, g);
    else if (c.nsec == 2) ax_headfilt_run<2, false>(w, cg, d, dr, c, st, g);
    else if (c.nsec == 4) ax_headfilt_run<4, false>(w, cg, d, dr, c, st, g);
    else ax_headfilt_run<5, false>(w, cg, d, dr, c, st, g);
}

// Bit edges of the head (greedy walk over its crossings, demodulate.py:85-93), then join the
// precomputed continuous crossings.
AX_HDN inline void ax_headwalk_item(const AxWave& w, int64_t cg) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (k >= dr.chunk_cap) return;
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk& ch = w.chunk[cg];
    const AxHeadGeom g = ax_head_geom(w, dr, st, c, ch, k);
    if (!g.active) return;
    const int64_t s = g.s, e = ch.e, len = g.len, H = g.H;
    int32_t* hz = w.head_idx + cg * (int64_t)w.head_zc_cap_max;
    float* ha1 = w.head_a1 + cg * (int64_t)w.head_zc_cap_max;
    float* ha2 = w.head_a2 + cg * (int64_t)w.head_zc_cap_max;
    ch.err = 0; ch.n_edges = 0; ch.n_head_edges = 0; ch.g_first = -1; ch.true_last = -1; ch.q_last = -1; ch.first_edge = -1;
    ch.merge_pos = -1; ch.n_pre = 0;
    const int nh = w.head_cnt[cg];
    if (nh < 0) { ch.err = AXCTD_DROP_CAPACITY; return; }
    const int32_t* zi = w.zc_idx + dr.zc_base;
    const uint8_t* nx = w.zc_nx + dr.zc_base;
    const uint64_t* cmask = w.cmask + dr.tile_base;
    const int32_t* crank = w.crank + dr.tile_base;
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
        hz[nhe] = hz[cpos]; ha1[nhe] = ha1[cpos]; ha2[nhe] = ha2[cpos];   // bit edge inside the head
        ++nhe;
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
        int64_t steps, merge;
        const int64_t endpos = ax_walk_end(nx, cmask, crank, pos, q, &steps, &merge);
        nedges += steps + 1;
        last = zi[endpos];
        // edges pos .. (explicit steps) .. merge .. (canonical crossings) .. endpos
        ch.merge_pos = merge;
        ch.n_pre = (int32_t)(merge < 0 ? steps + 1 : steps - (ax_canon_rank(cmask, crank, endpos) - ax_canon_rank(cmask, crank, merge)));
    }
    ch.n_head_edges = nhe;
    ch.n_edges = (int32_t)nedges;
    ch.true_last = last;
    // demodulate.py:100-101: a window that runs past the chunk end cannot be summed (only the final edge needs none)
    for (int t = 0; t < nhe; ++t)
        if (hz[t] + c.inset + c.npcm > len && t != nedges - 1) ch.err = AXCTD_DROP_SHORT_WINDOW;
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
            else if (w.streaming == 1) { st.n_chunks = k; st.chain_from = k; st.chain_end = 1; return; }   // the file may end here: decided by a later run
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

// ------------------------------------------------------------------ guard-band samples, settled exactly
// One listed sample (ax_unc_push) after the chunk chain is final.  The reference signs the output of
// scipy.signal.sosfilt run from zero state over the slice of ITS iteration (demodulate.py:74-78), so for every
// demodulated iteration that takes a crossing at this sample -- from the continuous pass past the head, or from the
// head recomputation of exactly that iteration -- the cascade is run again from the iteration's start in scipy's
// operation order without contraction (ax_biquad_exact) on the reference's own normalised samples
// (AXCTDprocessor.py:57), and its sign is compared with the one the fast pass used.  Samples outside every
// demodulated iteration (the lead-in before the first pulse, recordings without a pulse) do not matter.
AX_HD bool ax_exact_sign_neg(const AxWave& w, const AxDrop& dr, const AxCfg& c, const AxState& st, int64_t s, int64_t n) {
    const AxSrc x = ax_src(w, dr);
    double z[AX_MAXSEC][2];
    for (int q = 0; q < AX_MAXSEC; ++q) { z[q][0] = 0.0; z[q][1] = 0.0; }
    double y = 0.0;
    for (int64_t m = s; m <= n; ++m) {
        double u = x.xf ? x.xf[m] : ax_div(ax_sub((double)x.x[m], st.dc), st.ampl_d);
        for (int q = 0; q < c.nsec; ++q) u = ax_biquad_exact(u, c.sos[q], z[q][0], z[q][1]);
        y = u;
    }
    return y < 0.0;                                     // np.sign(y), zeros counted as positive (demodulate.py:77-78)
}
AX_HDN inline void ax_unc_resolve_item(const AxWave& w, int64_t item) {
    const int d = (int)(item / AX_UNC_CAP), j = (int)(item % AX_UNC_CAP);
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || j >= st.n_unc_listed) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    const int64_t* e = w.unc_list + ((int64_t)d * AX_UNC_CAP + j) * 2;
    const int64_t n = e[0] & 0xffffffffll;
    const bool neg = (e[0] >> 32) & 1;
    const int kp1 = (int)(e[0] >> 33);
    const AxChunk* ch = w.chunk + dr.chunk_base;
    bool relevant = false, ok = true;
    if (kp1 > 0) {                                      // head of iteration kp1 - 1, filtered from the start it had then
        const int k = kp1 - 1;
        if (k >= st.k0 && k < st.n_chunks && ch[k].s == e[1]) { relevant = true; ok = ax_exact_sign_neg(w, dr, c, st, ch[k].s, n) == neg; }
    } else {
        for (int k = st.k0; k < st.n_chunks && k < dr.chunk_cap; ++k) {
            if (ch[k].s > n) break;
            const int64_t len = ch[k].e - ch[k].s;
            const int64_t H = (w.force_exact || c.head > len) ? len : c.head;
            if (H >= len || n < ch[k].s + H - 1 || n > ch[k].e - 1) continue;      // crossings H-1 .. len-2 come from the continuous pass
            relevant = true;
            if (ax_exact_sign_neg(w, dr, c, st, ch[k].s, n) != neg) ok = false;
        }
    }
    if (relevant) {
        AX_ATOMIC_ADD32(&st.n_unc_relevant, 1);
        if (ok) AX_ATOMIC_ADD32(&st.n_unc_resolved, 1); else AX_ATOMIC_ADD32(&st.n_uncertain, 1);
    }
}
// More guard-band samples than the list holds: the unlisted ones are only known as counts per segment / per head;
// any that can lie inside a demodulated iteration stays unconfirmed.
AX_HDN inline void ax_unc_fin_item(const AxWave& w, int64_t d) {
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || st.n_unc_listed <= AX_UNC_CAP) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    const AxChunk* ch = w.chunk + dr.chunk_base;
    const int64_t lo = ch[st.k0].s + c.pad;
    int64_t cnt = 0;
    for (int j = 0; j < dr.nseg; ++j)
        if (((int64_t)j + 1) * w.seg_len > lo) cnt += w.seg_unc[dr.seg_base + j];
    for (int k = st.k0; k < st.n_chunks && k < dr.chunk_cap; ++k) cnt += w.head_unc[dr.chunk_base + k];
    const int64_t unlisted = cnt - st.n_unc_relevant;
    if (unlisted > 0) st.n_uncertain += (int32_t)(unlisted > 0x3fffffff ? 0x3fffffff : unlisted);
}

// ------------------------------------------------------------------ /2 decimation (AXCTDprocessor.py:60-62)
// scipy.signal.decimate(pcm, 2) = sosfiltfilt(cheby1(8, 0.05, 0.4, 'sos'), pcm)[::2]: odd extension by padlen
// samples, forward pass from the state zi*ext[0], backward pass over the reversed forward output from
// zi*(its first value), padding stripped, every second sample kept.  Both passes are cut into segments
// that start `dwarm` samples early from zero state (the recursion forgets its start like r^n; the very
// first segment of each pass starts from scipy's exact initial state).
AX_HD double ax_decim_ext(const AxWave& w, const AxDrop& dr, const AxCfg& c, const AxState& st, int64_t e) {
    const int16_t* x = w.pcm + dr.pcm_off;
    const int64_t N = dr.n_raw, P = c.dpad;
    // (x - dc) / ampl of AXCTDprocessor.py:57 as one FMA with the reciprocal, the form the staged kernel uses as well
#define AX_U(n) ax_fma((double)x[(n)], st.inv_ampl, -ax_mul(st.dc, st.inv_ampl))
    if (e < P) return ax_sub(ax_mul(2.0, AX_U(0)), AX_U(P - e));                 // 2 x[0] - x[P-e]
    if (e < P + N) return AX_U(e - P);
    return ax_sub(ax_mul(2.0, AX_U(N - 1)), AX_U(N - 2 - (e - P - N)));         // 2 x[N-1] - x[N-2-i]
#undef AX_U
}
AX_HD double ax_decim_step(const AxCfg& c, double u, double (*z)[2]) {
    for (int q = 0; q < c.dnsec; ++q) {
        const double* k = c.dsos[q];
        const double y = ax_fma(k[0], u, z[q][0]);
        z[q][0] = ax_fma(k[1], u, ax_fma(-k[4], y, z[q][1]));
        z[q][1] = ax_fma(k[2], u, -(k[5] * y));
        u = y;
    }
    return u;
}
// pass 0: forward over the extended recording -> w.fwd; pass 1: backward over w.fwd -> w.xf (even samples)
AX_HDN inline void ax_decim_item(const AxWave& w, int64_t sg, int pass) {
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::dseg_base, sg);
    const AxDrop& dr = w.drop[d];
    const int64_t j = sg - dr.dseg_base;
    if (dr.xf_off < 0 || j >= dr.ndseg) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const AxState& st = w.st[d];
    const int64_t E = dr.n_raw + 2 * (int64_t)c.dpad;
    const int64_t e0 = j * w.dseg_len;
    int64_t e1 = e0 + w.dseg_len; if (e1 > E) e1 = E;
    int64_t start = e0 - c.dwarm; if (start < 0) start = 0;
    double* fwd = w.fwd + dr.fwd_off;
    double z[AX_MAXSEC][2];
    for (int q = 0; q < AX_MAXSEC; ++q) { z[q][0] = 0.0; z[q][1] = 0.0; }
    if (start == 0) {
        const double x0 = pass == 0 ? ax_decim_ext(w, dr, c, st, 0) : fwd[E - 1];
        for (int q = 0; q < c.dnsec; ++q) { z[q][0] = ax_mul(c.dzi[q][0], x0); z[q][1] = ax_mul(c.dzi[q][1], x0); }
    }
    if (pass == 0) {
        for (int64_t e = start; e < e1; ++e) {
            const double y = ax_decim_step(c, ax_decim_ext(w, dr, c, st, e), z);
            if (e >= e0) fwd[e] = y;
        }
    } else {
        double* xf = w.xf + dr.xf_off;
        for (int64_t e = start; e < e1; ++e) {
            const double y = ax_decim_step(c, fwd[E - 1 - e], z);
            const int64_t m = (E - 1 - e) - c.dpad;                   // index in the unpadded recording
            if (e >= e0 && m >= 0 && m < dr.n_raw && (m & 1) == 0) xf[m >> 1] = y;
        }
    }
}
// downstream kernels see an already normalised signal
AX_HDN inline void ax_decim_fin(const AxWave& w, int64_t d) {
    if (w.drop[d].xf_off < 0) return;
    AxState& st = w.st[d];
    st.dc_raw = st.dc; st.ampl_raw = st.ampl_d;
    st.dc = 0.0; st.inv_ampl = 1.0; st.ampl_d = 1.0;
}
