// ax_kernels.cuh -- CUDA-only cooperative kernels (sm_100a).
//
//   k_stats_coalesced   int16 sum / max|x| with 128-bit loads (AXCTDprocessor.py:55-56)
//   k_tone_blocks       400 / 7500 / dead-frequency single-bin DFTs by gcd(N_power, d_pcm)
//                       blocks with the cos/sin table staged in shared memory
//                       (AXCTDprocessor.py:358-364)
//   k_tone_combine      5 rotated block sums -> one 0.1 s window magnitude
#pragma once
#include <cuda_runtime.h>
#include "ax_proto.h"

// ------------------------------------------------------------------ stats
__global__ void __launch_bounds__(256) k_stats_coalesced(AxWave w) {
    const int64_t slab = blockIdx.x;
    const int d = w.slab_drop[slab];
    const AxDrop& dr = w.drop[d];
    const int64_t j = slab - dr.slab_base;
    const int64_t a = j * AX_STAT_SLAB;
    int64_t b = a + AX_STAT_SLAB;
    if (b > dr.n) b = dr.n;
    const int16_t* x = w.pcm + dr.pcm_off + a;          // 128-byte aligned
    const int cnt = (int)(b - a);
    const int nvec = cnt >> 3;
    long long sum = 0;
    int mx = -32768;
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const uint4 q = __ldg(xv + v);
        const unsigned wds[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int lo = (short)(wds[t] & 0xFFFFu), hi = (short)(wds[t] >> 16);
            sum += lo + hi;
            const int alo = (lo == -32768) ? -32768 : abs(lo), ahi = (hi == -32768) ? -32768 : abs(hi);
            mx = max(mx, max(alo, ahi));
        }
    }
    for (int t = (nvec << 3) + threadIdx.x; t < cnt; t += blockDim.x) {
        const int v = x[t];
        sum += v;
        mx = max(mx, (v == -32768) ? -32768 : abs(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ long long ssum[8];
    __shared__ int smx[8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { ssum[wid] = sum; smx[wid] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 1; q < 8; ++q) { sum += ssum[q]; mx = max(mx, smx[q]); }
        atomicAdd((unsigned long long*)&w.st[d].sum, (unsigned long long)sum);
        atomicMax(&w.st[d].ampl, mx);
    }
}

static inline void ax_launch_stats(const AxWave& w, cudaStream_t stream) {
    if (w.nslab_total > 0) k_stats_coalesced<<<w.nslab_total, 256, 0, stream>>>(w);
}

// ------------------------------------------------------------------ fused demodulation pass
// k_demod_fused: int16 PCM -> normalise -> Butterworth SOS cascade (double) -> zero crossings
// (demodulate.py:74-79) -> mark / space window magnitudes after every crossing (demodulate.py:99-102).
//
// Work decomposition: one lane per segment of seg_len samples (plus the warm-up overlap), 64 samples
// ("a row") per iteration; a warp therefore advances 32 independent filter recurrences in lockstep.
//   staging  each warp copies one 128-byte line per lane per stage with 16-byte cp.async (LDGSTS),
//            double buffered, into lane-major rows that the owner reads back with conflict-free LDS.128;
//   phase 1  the owner lane runs the cascade over its 64 samples (13 DFMA-pipe operations per sample for
//            three sections), collects the sign bits in registers and stores the float roundings of y
//            into its 128-sample ring in shared memory (STS.128, conflict free);
//   phase 2  the crossings of the previous row are compacted across the warp and dealt to the lanes one
//            each, so the 4*NPCM-FMA fp32 windows (ax_window32, phasors as constant-bank operands) run
//            without divergence whatever the crossing density of the individual rows.
// y never leaves the SM; the only global traffic is the int16 stream in and (index, |S1|, |S2|) per
// crossing out.
#define AX_FD_WARPS 4
#define AX_FD_THREADS (AX_FD_WARPS * 32)
#define AX_FD_ROW 72                                   // int16 per staged row: 64 samples + 8 pad (144-byte stride)
#define AX_FD_STAGE (32 * AX_FD_ROW)                   // int16 per warp per stage
#define AX_FD_YSTRIDE 132                              // floats per lane ring: 128 + 4 pad (33 quads: conflict-free STS.128)
#define AX_FD_LIST 256                                 // crossings dealt per pass

struct AxFdSmem {
    int16_t stage[2][AX_FD_STAGE];
    float yring[32 * AX_FD_YSTRIDE];
    uint32_t list[AX_FD_LIST];
    int32_t row_begin[32], row_stop[32];
};

__device__ __forceinline__ void ax_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void ax_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void ax_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NSEC, int NPCM>
__global__ void __launch_bounds__(AX_FD_THREADS, 2) k_demod_fused(const __grid_constant__ AxWave w, const __grid_constant__ AxWinTab tab, int cfg_id) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];
    const int d = w.seg_drop[(int64_t)blockIdx.x * AX_FD_THREADS];
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;                       // another launch handles this rate class
    const AxCfg& c = w.cfg[cfg_id];
    AxState& st = w.st[d];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AxFdSmem& sm = reinterpret_cast<AxFdSmem*>(ax_smem_raw)[warp];
    const int64_t seg = (int64_t)blockIdx.x * AX_FD_THREADS + threadIdx.x;
    const int64_t j = seg - dr.seg_base;
    const bool active = j < dr.nseg;
    AxSegGeom g;
    g.seg_start = g.seg_end = g.n_begin = g.n_stop = 0;
    if (active) g = ax_seg_geom(dr, c, w.seg_len, j);
    const int T = active ? (int)((g.n_stop - g.n_begin + 63) >> 6) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    const int16_t* xdrop = w.pcm + dr.pcm_off;
    const unsigned long long xrow = (unsigned long long)(xdrop + g.n_begin);      // 16-byte aligned
    sm.row_begin[lane] = (int)g.n_begin; sm.row_stop[lane] = (int)g.n_stop;
    // ---- cascade constants (Butterworth form, see AxFilt::filter)
    double z0[NSEC], z1[NSEC], a1[NSEC], a2[NSEC], sg[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        z0[s] = 0.0; z1[s] = 0.0;
        a1[s] = -c.sos[s][4]; a2[s] = -c.sos[s][5];
        sg[s] = (c.sos[s][1] < 0.0) ? -2.0 : 2.0;
    }
    const double k0 = c.sos[0][0] * st.inv_ampl, k1 = c.sos[0][0] * -(st.dc * st.inv_ampl);
    const unsigned guard_hi = (unsigned)__double2hiint(w.guard);
    const int nb = (int)g.n_begin, nstop = (int)g.n_stop, sstart = (int)g.seg_start, send = (int)g.seg_end;
    const int64_t slot0 = seg * (int64_t)w.seg_cap;
    const int64_t wslot0 = (seg - lane) * (int64_t)w.seg_cap;       // slot of the warp's row 0
    // rows this lane helps to stage: r = i*4 + prow, i = 0..7
    const int prow = lane >> 3, piece = lane & 7;
    unsigned long long src[8];
    int Tr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        src[i] = __shfl_sync(0xffffffffu, xrow, i * 4 + prow) + (unsigned long long)piece * 16;
        Tr[i] = __shfl_sync(0xffffffffu, T, i * 4 + prow);
    }
    auto issue = [&](int t, int s) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (t < Tr[i]) ax_cp_async16(&sm.stage[s][(i * 4 + prow) * AX_FD_ROW + piece * 8],
                                         reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
        ax_cp_async_commit();
    };
    unsigned long long Sprev = 0ull;        // sign bits of row t-1 (bit i = sample i negative)
    int count = 0, unc = 0;
    float* myring = sm.yring + lane * AX_FD_YSTRIDE;
    __syncwarp();
    if (Tmax > 0) issue(0, 0);
    for (int t = 0; t <= Tmax; ++t) {
        unsigned long long Scur = 0ull;
        if (t < Tmax) {
            if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
            __syncwarp();
            // ---------------- phase 1: the cascade over this lane's 64 samples
            if (t < T) {
                const int4* rp = reinterpret_cast<const int4*>(&sm.stage[t & 1][lane * AX_FD_ROW]);
                float4* yo = reinterpret_cast<float4*>(myring + (t & 1) * 64);
                unsigned minabs = 0x7fffffffu;
#pragma unroll
                for (int hw = 0; hw < 2; ++hw) {
                    unsigned sb = 0u;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int4 q = rp[hw * 4 + v];
                        const int wd[4] = {q.x, q.y, q.z, q.w};
                        float yf[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int xi = (e & 1) ? (wd[e >> 1] >> 16) : (int)(short)(wd[e >> 1] & 0xFFFF);
                            double tt = fma((double)xi, k0, k1);
#pragma unroll
                            for (int s = 0; s < NSEC; ++s) {
                                const double y = tt + z0[s];
                                z0[s] = fma(a1[s], y, fma(sg[s], tt, z1[s]));
                                z1[s] = fma(a2[s], y, tt);
                                tt = y;
                            }
                            const unsigned hi = (unsigned)__double2hiint(tt);
                            sb = __funnelshift_l(hi, sb, 1);                 // MSB-first: sample 0 of this half ends at bit 31
                            minabs = min(minabs, hi & 0x7fffffffu);
                            yf[e] = (float)tt;
                        }
                        yo[hw * 8 + v * 2] = make_float4(yf[0], yf[1], yf[2], yf[3]);
                        yo[hw * 8 + v * 2 + 1] = make_float4(yf[4], yf[5], yf[6], yf[7]);
                    }
                    Scur |= (unsigned long long)__brev(sb) << (32 * hw);
                }
                // guard band: a filter output this close to zero cannot be signed reliably (AXCTD_DROP_UNCERTAIN)
                if (minabs < guard_hi) {
                    const int base = nb + 64 * t;
                    for (int i = 0; i < 64; ++i) {
                        const int n = base + i;
                        if (n >= sstart && n < send && n < nstop && fabsf(myring[(t & 1) * 64 + i]) < (float)w.guard) ++unc;
                    }
                }
            }
        }
        __syncwarp();
        // ---------------- phase 2: crossings of row t-1 and their windows
        if (t >= 1) {
            unsigned long long X = 0ull;
            const int base = nb + 64 * (t - 1);
            if (t - 1 < T) {
                const unsigned long long nxt = (t < T) ? (Scur & 1ull) : ((Sprev >> 63) & 1ull);
                X = Sprev ^ ((Sprev >> 1) | (nxt << 63));
                int lo = sstart - base, hi = min(send, nstop - 1) - base;      // crossing i needs sample i+1
                lo = max(lo, 0); hi = min(hi, 64);
                unsigned long long m = 0ull;
                if (hi > lo) m = ((hi >= 64) ? ~0ull : ((1ull << hi) - 1ull)) & ~((1ull << lo) - 1ull);
                X &= m;
            }
            while (__any_sync(0xffffffffu, X != 0ull)) {
                const int take = min(__popcll(X), AX_FD_LIST / 32);
                int off = take;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
                const int total = __shfl_sync(0xffffffffu, off, 31);
                off -= take;
                for (int q = 0; q < take; ++q) {
                    const int p = __ffsll((long long)X) - 1;
                    X &= X - 1ull;
                    sm.list[off + q] = (unsigned)p | ((unsigned)lane << 6) | ((unsigned)(count + q) << 11);
                }
                count += take;
                __syncwarp();
                for (int it = lane; it < total; it += 32) {
                    const unsigned en = sm.list[it];
                    const int p = (int)(en & 63u), r = (int)((en >> 6) & 31u), op = (int)(en >> 11);
                    const int j0 = ((t - 1) & 1) * 64 + p + 1;           // ring position of the first window sample
                    const int o = j0 & 3;
                    const float4* rq = reinterpret_cast<const float4*>(sm.yring + r * AX_FD_YSTRIDE);
                    constexpr int NQ = (NPCM + 6) >> 2;
                    float yv[NQ * 4];
#pragma unroll
                    for (int k = 0; k < NQ; ++k) {
                        const float4 v = rq[((j0 >> 2) + k) & 31];
                        yv[4 * k] = v.x; yv[4 * k + 1] = v.y; yv[4 * k + 2] = v.z; yv[4 * k + 3] = v.w;
                    }
                    float m1, m2;
                    ax_window32(yv, o, NPCM, tab, &m1, &m2);
                    const int rb = sm.row_begin[r] + 64 * (t - 1);
                    const bool complete = rb + p + NPCM < sm.row_stop[r];
                    if (op < w.seg_cap) {
                        const int64_t oi = wslot0 + (int64_t)r * w.seg_cap + op;
                        w.rec_idx[oi] = rb + p;
                        w.rec_a1[oi] = complete ? m1 : __int_as_float(0x7fc00000);
                        w.rec_a2[oi] = complete ? m2 : __int_as_float(0x7fc00000);
                    }
                }
                __syncwarp();
            }
        }
        Sprev = Scur;
        __syncwarp();
    }
    if (active) {
        if (count > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; count = w.seg_cap; }
        w.seg_cnt[seg] = count;
        if (unc) atomicAdd(&st.n_uncertain, unc);
    } else w.seg_cnt[seg] = 0;
    (void)slot0;
}

template <int NSEC, int NPCM>
static inline void ax_launch_demod_fused(const AxWave& w, const AxCfg& c, int cfg_id, cudaStream_t stream) {
    const size_t smem = AX_FD_WARPS * sizeof(AxFdSmem);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_demod_fused<NSEC, NPCM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    k_demod_fused<NSEC, NPCM><<<w.nseg_total / AX_FD_THREADS, AX_FD_THREADS, smem, stream>>>(w, c.win_tab, cfg_id);
}

// true if the fused kernel has an instantiation for this rate class
static inline bool ax_demod_fused_ok(const AxCfg& c) {
    return ax_sos_is_butter(c) && (c.nsec == 3 || c.nsec == 6) && (c.npcm == 39 || c.npcm == 43) && c.inset == 1;
}
static inline void ax_launch_demod_fused_any(const AxWave& w, const AxCfg& c, int cfg_id, cudaStream_t stream) {
    if (c.nsec == 3 && c.npcm == 39) ax_launch_demod_fused<3, 39>(w, c, cfg_id, stream);
    else if (c.nsec == 3 && c.npcm == 43) ax_launch_demod_fused<3, 43>(w, c, cfg_id, stream);
    else if (c.nsec == 6 && c.npcm == 39) ax_launch_demod_fused<6, 39>(w, c, cfg_id, stream);
    else ax_launch_demod_fused<6, 43>(w, c, cfg_id, stream);
}

// ------------------------------------------------------------------ bit decisions with shared window sums
// As ax_bits_item, one CTA per run() iteration (no per-bit searches); a window that needs double
// precision is summed by the whole warp.
__global__ void __launch_bounds__(128) k_bits_chunk(AxWave w, int phase) {
    const int64_t cg = blockIdx.x;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.sm_status < 1 || st.nedges_total == 0 || k < st.k0 || k >= st.n_chunks || k >= dr.chunk_cap) return;
    const AxChunk& ch = w.chunk[cg];
    const int nb = ch.n_edges - 1;
    if (nb <= 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t slot0 = dr.edge_base + ch.bit_off;
    for (int jb = threadIdx.x - lane; jb < nb; jb += blockDim.x) {       // warp-uniform trip count
        const int mine = jb + lane;
        const int64_t slot = slot0 + mine;
        AxBitFix fx;
        fx.d = 0; fx.i = 0; fx.q0 = 0;
        const bool need = mine < nb && ax_bits_need(w, d, k, slot, phase, &fx);
        unsigned ball = __ballot_sync(0xffffffffu, need);
        while (ball) {
            const int L = __ffs((int)ball) - 1;
            ball &= ball - 1;
            AxBitFix f;
            f.d = d;
            f.i = __shfl_sync(0xffffffffu, fx.i, L);
            f.q0 = __shfl_sync(0xffffffffu, fx.q0, L);
            double acc[4];
            ax_gwin_partial(w.pcm + dr.pcm_off, f.i, f.q0, w.cfg[dr.cfg], lane, 32, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            if (lane == L) ax_bits_fix(w, slot, f, acc);
        }
        if (phase == 1 && mine < nb) ax_bits_decide(w, d, slot);
    }
}

// ------------------------------------------------------------------ bit edges (CTA per run() iteration)
// ax_emit_item with one thread per edge: the continuous part of a chunk's walk is read off the
// canonical walk by rank (ax_emit_canon_pos); only the few edges stepped explicitly before the walk
// joins it are produced by one thread.
__global__ void __launch_bounds__(128) k_emit_chunk(AxWave w) {
    const int64_t cg = blockIdx.x;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (!ax_emit_active(w, dr, st, k)) return;
    AxChunk& ch = w.chunk[cg];
    const int ne = ch.n_edges;
    if (ne <= 0) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int nhe = ch.n_head_edges, npre = ch.n_pre;
    const bool merged = ch.merge_pos >= 0;
    for (int t = threadIdx.x; t < ne; t += blockDim.x) {
        if (t < nhe) ax_emit_edge(w, dr, st, c, ch, cg, k, t, 0);
        else if (t >= nhe + npre) { if (merged) ax_emit_edge(w, dr, st, c, ch, cg, k, t, ax_emit_canon_pos(w, dr, ch, t)); }
        else if (t == nhe) {
            const uint8_t* nx = w.zc_nx + dr.zc_base;
            int64_t pos = ch.g_first;
            for (int q = 0; q < npre; ++q) { ax_emit_edge(w, dr, st, c, ch, cg, k, nhe + q, pos); if (q < npre - 1) pos += nx[pos]; }
        }
    }
}

// ------------------------------------------------------------------ dense crossing arrays
// ax_compact_item with a warp per segment (coalesced), and the walk steps (ax_nx_item) with a 2-D grid.
__global__ void __launch_bounds__(256) k_compact_warp(AxWave w) {
    const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (seg >= w.nseg_total) return;
    const int lane = threadIdx.x & 31;
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    if (w.st[d].zc_count == 0) return;
    const int64_t src = seg * (int64_t)w.seg_cap, dst = dr.zc_base + w.seg_off[seg] + w.blk_sum[seg / 128];
    const int cnt = w.seg_cnt[seg];
    for (int q = lane; q < cnt; q += 32) {
        w.zc_idx[dst + q] = w.rec_idx[src + q];
        w.zc_a1[dst + q] = w.rec_a1[src + q];
        w.zc_a2[dst + q] = w.rec_a2[src + q];
    }
}

__global__ void __launch_bounds__(256) k_nx_grid(AxWave w) {
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    const int64_t M = w.st[d].zc_count;
    const AxCfg& c = w.cfg[dr.cfg];
    const int32_t* zi = w.zc_idx + dr.zc_base;
    for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < M; pos += (int64_t)gridDim.x * blockDim.x)
        w.zc_nx[dr.zc_base + pos] = (pos + 4 < M) ? (uint8_t)(ax_next(zi, pos, c.fs2, 2 * (int64_t)c.bitrate) - pos) : (uint8_t)0;
}

// ------------------------------------------------------------------ canonical walk tables (block per drop)
// Same result as ax_canon_item.  The per-tile exit maps compose associatively, so each thread folds a
// contiguous range of tiles into one 4-state map, the ranges are chained through shared memory, and a
// second sweep writes the visited masks and the running counts.
#define AX_CANON_THREADS 256
__global__ void __launch_bounds__(AX_CANON_THREADS) k_canon_block(AxWave w) {
    const int d = blockIdx.x;
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int64_t M = st.zc_count;
    const int ntile = (int)((M + AX_TILE - 1) / AX_TILE);
    uint64_t* cmask = w.cmask + dr.tile_base;
    int32_t* crank = w.crank + dr.tile_base;
    const uint64_t* tmask = w.tile_mask + (int64_t)dr.tile_base * 4;
    const uint32_t* tmap = w.tile_map + dr.tile_base;
    __shared__ int64_t s_entry;
    __shared__ int s_state[AX_CANON_THREADS + 1];
    __shared__ uint32_t s_map[AX_CANON_THREADS];
    __shared__ int s_cnt[AX_CANON_THREADS + 1];
    const int tid = threadIdx.x;
    if (tid == 0) {
        const int64_t entry = ax_lower_bound(w.zc_idx + dr.zc_base, M, w.chunk[dr.chunk_base + st.k0].s + c.pad);
        s_entry = entry;
        if (entry < M) {
            uint64_t m0;
            s_state[0] = ax_canon_first_tile(w.zc_nx + dr.zc_base, M, entry, &m0);
            const int t0 = (int)(entry / AX_TILE);
            cmask[t0] = m0; crank[t0] = 0;
            s_cnt[0] = ax_popc64(m0);
        }
    }
    __syncthreads();
    const int64_t entry = s_entry;
    const int t0 = (int)(entry / AX_TILE);
    for (int t = tid; t < ntile && t < t0; t += AX_CANON_THREADS) { cmask[t] = 0; crank[t] = 0; }
    if (entry >= M) return;
    const int nrest = ntile - (t0 + 1);
    const int per = (nrest + AX_CANON_THREADS - 1) / AX_CANON_THREADS;
    const int ta = min(t0 + 1 + tid * per, ntile), tb = min(ta + per, ntile);
    {   // fold my range into one map
        uint32_t f = 0x03020100u;
        for (int t = ta; t < tb; ++t) {
            const uint32_t m = tmap[t];
            uint32_t g = 0;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const uint32_t sI = (f >> (8 * o)) & 0xFFu;
                const uint32_t e = sI < 4u ? ((m >> (8 * sI)) & 0xFFu) : 0xFFu;
                g |= e << (8 * o);
            }
            f = g;
        }
        s_map[tid] = f;
    }
    __syncthreads();
    if (tid == 0) {
        int state = s_state[0];
        for (int i = 0; i < AX_CANON_THREADS; ++i) {
            s_state[i] = state;
            state = ax_map_apply(s_map[i], state);
        }
    }
    __syncthreads();
    int state = s_state[tid];
    int local = 0;
    for (int t = ta; t < tb; ++t) {
        const uint64_t m = state < 4 ? tmask[(int64_t)t * 4 + state] : 0ull;
        cmask[t] = m; crank[t] = local;
        local += __popcll(m);
        state = ax_map_apply(tmap[t], state);
    }
    const int first_cnt = s_cnt[0];
    __syncthreads();
    s_cnt[tid + 1] = local;
    __syncthreads();
    if (tid == 0) {
        int run = first_cnt;
        for (int i = 0; i < AX_CANON_THREADS; ++i) { const int v = s_cnt[i + 1]; s_cnt[i + 1] = run; run += v; }
    }
    __syncthreads();
    const int base = s_cnt[tid + 1];
    for (int t = ta; t < tb; ++t) crank[t] += base;
}

// ------------------------------------------------------------------ tones
#define AX_TONE_WARPS 8
#define AX_TONE_R 4          // blocks per warp pass (register blocking against the smem table)

__device__ __forceinline__ bool ax_tone_chunk_active(const AxWave& w, const AxState& st, int k, int phase_b) {
    int klo, khi;
    return ax_level_range(w, st, phase_b, &klo, &khi) && k >= klo && k < khi;
}

__global__ void __launch_bounds__(AX_TONE_WARPS * 32)
k_tone_blocks(AxWave w, int cfg_id, int phase_b, int qpc, int chunk_total) {
    extern __shared__ double tab[];                      // [6][G]
    const AxCfg& c = w.cfg[cfg_id];
    const int G = c.tone_G;
    for (int i = threadIdx.x; i < 6 * G; i += blockDim.x) {
        const int q = i / G, m = i - q * G;
        tab[i] = c.tone_cs[6 * (int64_t)m + q];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * AX_TONE_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * AX_TONE_WARPS;
    const int64_t total = (int64_t)chunk_total * qpc;
    for (int64_t item = gw; item < total; item += nw) {
        const int64_t cg = item / qpc;
        const int quad = (int)(item - cg * qpc);
        const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
        const AxDrop& dr = w.drop[d];
        if (dr.cfg != cfg_id) continue;
        const AxState& st = w.st[d];
        const int k = (int)(cg - dr.chunk_base);
        if (k >= dr.chunk_cap || st.status >= AXCTD_DROP_CAPACITY || !ax_tone_chunk_active(w, st, k, phase_b)) continue;
        const AxChunk& ch = w.chunk[cg];
        if (ch.np <= 0) continue;
        const int B = (ch.np - 1) * c.tone_stride + c.tone_nb;
        const int b0 = quad * AX_TONE_R;
        if (b0 >= B) continue;
        const int16_t* x = w.pcm + dr.pcm_off + ch.s + (int64_t)b0 * G;
        const double kmul = st.inv_ampl, kadd = -(st.dc * st.inv_ampl);
        double acc[AX_TONE_R][6];
#pragma unroll
        for (int r = 0; r < AX_TONE_R; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[r][q] = 0.0;
        const int nblk = min(AX_TONE_R, B - b0);
        for (int m = lane; m < G; m += 32) {
            double t[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) t[q] = tab[q * G + m];
#pragma unroll
            for (int r = 0; r < AX_TONE_R; ++r) {
                if (r < nblk) {
                    const double u = fma((double)x[(int64_t)r * G + m], kmul, kadd);
#pragma unroll
                    for (int q = 0; q < 6; ++q) acc[r][q] = fma(u, t[q], acc[r][q]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < AX_TONE_R; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                double v = acc[r][q];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                acc[r][q] = v;
            }
        if (lane == 0) {
            double* out = w.blk + ((int64_t)cg * w.blk_stride + b0) * 6;
            for (int r = 0; r < nblk; ++r)
                for (int q = 0; q < 6; ++q) out[r * 6 + q] = acc[r][q];
        }
    }
}

__global__ void k_tone_combine(AxWave w, int cfg_id, int phase_b) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= w.pw_total) return;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::pw_base, slot);
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;
    const AxState& st = w.st[d];
    if (st.status >= AXCTD_DROP_CAPACITY) return;
    const AxCfg& c = w.cfg[cfg_id];
    const int32_t i = (int32_t)(slot - dr.pw_base);
    const AxChunk* ch = w.chunk + dr.chunk_base;
    int klo, khi;                                         // active chunk range [klo, khi)
    if (!ax_level_range(w, st, phase_b, &klo, &khi)) return;
    if (i < ch[klo].pw_off || i >= ch[khi - 1].pw_off + ch[khi - 1].np) return;
    int lo = klo, hi = khi - 1;                           // last chunk with pw_off <= i
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ch[mid].pw_off <= i) lo = mid; else hi = mid - 1; }
    const int k = lo;
    const int jw = i - ch[k].pw_off;
    if (jw >= ch[k].np) return;
    const double* S = w.blk + ((int64_t)(dr.chunk_base + k) * w.blk_stride + (int64_t)jw * c.tone_stride) * 6;
    double re[3] = {0, 0, 0}, im[3] = {0, 0, 0};
    for (int q = 0; q < c.tone_nb; ++q) {
        const double* rot = c.tone_cs + 6 * (int64_t)q * c.tone_G;     // e^{j theta_f G q}
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const double sr = S[q * 6 + 2 * f], si = S[q * 6 + 2 * f + 1];
            const double cr = rot[2 * f], sn = rot[2 * f + 1];
            re[f] += sr * cr - si * sn;
            im[f] += sr * sn + si * cr;
        }
    }
#pragma unroll
    for (int f = 0; f < 3; ++f) w.pw_raw[f * (int64_t)w.pw_total + slot] = hypot(re[f], im[f]);
}

static inline void ax_launch_tone_blocked(const AxWave& w, int cfg_id, const AxCfg& c, int phase_b, int chunk_total,
                                          cudaStream_t stream) {
    const int qpc = (w.blk_stride + AX_TONE_R - 1) / AX_TONE_R;
    const int64_t total = (int64_t)chunk_total * qpc;
    if (total <= 0 || w.pw_total <= 0) return;
    const size_t smem = (size_t)6 * c.tone_G * sizeof(double);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_tone_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
    int grid = (int)std::min<int64_t>((total + AX_TONE_WARPS - 1) / AX_TONE_WARPS, 148 * 4);
    k_tone_blocks<<<grid, AX_TONE_WARPS * 32, smem, stream>>>(w, cfg_id, phase_b, qpc, chunk_total);
    k_tone_combine<<<(w.pw_total + 127) / 128, 128, 0, stream>>>(w, cfg_id, phase_b);
}

