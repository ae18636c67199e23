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

// ------------------------------------------------------------------ filter (staged)
// One thread per segment, as ax_filter_segment, but the int16 samples reach the
// threads through shared memory: each warp copies, with 16-byte cp.async, one
// full 128-byte line (64 samples) per thread per stage, double buffered, and
// every thread then reads its own row with conflict-free 128-bit LDS.  The
// cos/sin table of the bit windows sits in shared memory as well (one config
// per CTA: segment ranges are padded to multiples of 128 per drop).
#define AX_FS_THREADS 128
#define AX_FS_ROW 72                                  // 64 samples + 8 pad (144-byte row stride)
#define AX_FS_STAGE (4 * 32 * AX_FS_ROW)              // int16 elements per stage

__device__ __forceinline__ void ax_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void ax_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void ax_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int NSEC, bool BUTTER>
__global__ void __launch_bounds__(AX_FS_THREADS) k_filter_staged(AxWave w) {
    extern __shared__ __align__(16) unsigned char ax_smem[];
    const int d = w.seg_drop[(int64_t)blockIdx.x * AX_FS_THREADS];
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    if (c.nsec != NSEC || ax_sos_is_butter(c) != BUTTER) return;      // another instantiation handles this drop
    AxState& st = w.st[d];
    const int R = c.rebase;
    double* tab = reinterpret_cast<double*>(ax_smem);                   // [R][4]
    int16_t* stage = reinterpret_cast<int16_t*>(ax_smem + (size_t)R * 4 * sizeof(double));
    for (int i = threadIdx.x; i < 4 * R; i += AX_FS_THREADS) tab[i] = c.bit_cs[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * AX_FS_THREADS + threadIdx.x;
    const int64_t j = seg - dr.seg_base;
    const bool active = j < dr.nseg;
    AxSegGeom g;
    g.seg_start = g.seg_end = g.n_begin = g.n_stop = 0;
    if (active) g = ax_seg_geom(dr, c, w.seg_len, j);
    const int T = active ? (int)((g.n_stop - g.n_begin + 63) >> 6) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    const unsigned long long xrow = (unsigned long long)(w.pcm + dr.pcm_off + g.n_begin);   // 16-byte aligned
    const int64_t slot = seg * (int64_t)w.seg_cap;
    AxFilt<NSEC, BUTTER> f;
    f.init(c, st, (int32_t)g.seg_start, (int32_t)g.seg_end, w.guard, w.rec_idx + slot, w.rec_a1 + slot, w.rec_a2 + slot, w.seg_cap);
    int16_t* wst = stage + warp * (32 * AX_FS_ROW);
    const int prow = lane >> 3, piece = lane & 7;
    // rows this lane helps to copy: r = i*4 + prow, i = 0..7
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
            if (t < Tr[i]) ax_cp_async16(wst + s * AX_FS_STAGE + (i * 4 + prow) * AX_FS_ROW + piece * 8,
                                         reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
        ax_cp_async_commit();
    };
    if (Tmax > 0) issue(0, 0);
    for (int t = 0; t < Tmax; ++t) {
        if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
        __syncwarp();
        if (t < T) {
            const int4* rp = reinterpret_cast<const int4*>(wst + (t & 1) * AX_FS_STAGE + lane * AX_FS_ROW);
            int32_t n = (int32_t)g.n_begin + t * 64;
            const int32_t n_stop = (int32_t)g.n_stop;
#pragma unroll 1
            for (int v = 0; v < 8; ++v) {
                const int4 q = rp[v];
                const int wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const int lo = (short)(wd[h] & 0xFFFF), hi = wd[h] >> 16;
                    if (n < n_stop) f.step(n, (double)lo, tab);
                    ++n;
                    if (n < n_stop) f.step(n, (double)hi, tab);
                    ++n;
                }
            }
        }
        __syncwarp();
    }
    if (!active) { w.seg_cnt[seg] = 0; return; }
    f.finish();
    if (f.cnt > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; f.cnt = w.seg_cap; }
    w.seg_cnt[seg] = f.cnt;
    if (f.unc) atomicAdd(&st.n_uncertain, f.unc);
}

template <int NSEC, bool BUTTER>
static inline void ax_launch_filter_variant(const AxWave& w, int rebase_max, cudaStream_t stream) {
    const size_t smem = (size_t)rebase_max * 4 * sizeof(double) + 2 * AX_FS_STAGE * sizeof(int16_t);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_filter_staged<NSEC, BUTTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr_set = true; }
    k_filter_staged<NSEC, BUTTER><<<w.nseg_total / AX_FS_THREADS, AX_FS_THREADS, smem, stream>>>(w);
}

// ------------------------------------------------------------------ filter (fp32, fused)
// The production filter pass.  B200's double-precision pipe is ~30x slower than the FP32 pipe,
// so the streaming work runs in fp32 and double precision is spent only where a decision could
// depend on it:
//   phase 1 (one thread per segment): int16 -> Butterworth SOS cascade in fp32 (13 FMA-pipe
//            operations per sample) -> y written to a per-thread row in shared memory;
//   phase 2a (warp-cooperative): every sample with |y| < guard32 is re-evaluated in fp64 as a
//            direct convolution with the cascade's impulse response (ax_fir_partial) and
//            replaced, so that all signs used below are the exact filter's signs;
//   phase 2b (warp-cooperative): zero crossings of the previous row (demodulate.py:77-79) and,
//            for each, the mark / space single-bin DFT magnitudes of the following npcm samples
//            (demodulate.py:99-102), 32 lanes over the window taps.
// The PCM reaches shared memory through 16-byte cp.async copies, one full 128-byte line per
// thread per stage, double buffered; y never leaves the SM.
#define AX_F32_YROW 65                                 // floats per y row: 64 + 1 pad (conflict-free column access)

template <int NSEC>
__global__ void __launch_bounds__(AX_FS_THREADS, 2) k_filter32(AxWave w) {
    extern __shared__ __align__(16) unsigned char ax_smem[];
    const int d = w.seg_drop[(int64_t)blockIdx.x * AX_FS_THREADS];
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    if (c.nsec != NSEC || !ax_sos_is_butter(c) || c.npcm > 64) return;
    AxState& st = w.st[d];
    float4* tabf = reinterpret_cast<float4*>(ax_smem);                             // [64] cos1,sin1,cos2,sin2
    int16_t* stage = reinterpret_cast<int16_t*>(ax_smem + 64 * sizeof(float4));   // [2][4][32][72]
    float* yrow = reinterpret_cast<float*>(ax_smem + 64 * sizeof(float4) + 2 * AX_FS_STAGE * sizeof(int16_t));   // [2][128][65]
    int* count = reinterpret_cast<int*>(yrow + 2 * AX_FS_THREADS * AX_F32_YROW);   // [128]
    const int npcm = c.npcm;
    if (threadIdx.x < 64) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((int)threadIdx.x < npcm) {
            const double* t4 = c.bit_cs + 4 * threadIdx.x;
            v = make_float4((float)t4[0], (float)t4[1], (float)t4[2], (float)t4[3]);
        }
        tabf[threadIdx.x] = v;
    }
    count[threadIdx.x] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t seg = (int64_t)blockIdx.x * AX_FS_THREADS + threadIdx.x;
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
    const unsigned long long xrow = (unsigned long long)(xdrop + g.n_begin);
    // ---- fp32 cascade state (Butterworth form, see AxFilt)
    float z0[NSEC], z1[NSEC], a1[NSEC], a2[NSEC], sg[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        z0[s] = 0.f; z1[s] = 0.f;
        a1[s] = (float)c.sos[s][4]; a2[s] = (float)c.sos[s][5];
        sg[s] = (c.sos[s][1] < 0.0) ? -2.f : 2.f;
    }
    const float k0 = (float)(c.sos[0][0] * st.inv_ampl), k1 = (float)(-(c.sos[0][0] * st.dc * st.inv_ampl));
    const float G = w.guard32;
    int16_t* wst = stage + warp * (32 * AX_FS_ROW);
    float* wy = yrow + (warp * 32) * AX_F32_YROW;          // this warp's 32 rows inside one y buffer
    const int prow = lane >> 3, piece = lane & 7;
    unsigned long long src[8];
    int Tr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        src[i] = __shfl_sync(0xffffffffu, xrow, i * 4 + prow) + (unsigned long long)piece * 16;
        Tr[i] = __shfl_sync(0xffffffffu, T, i * 4 + prow);
    }
    const int nb = (int)g.n_begin, nstop = (int)g.n_stop, sstart = (int)g.seg_start, send = (int)g.seg_end;
    const int64_t slot0 = seg * (int64_t)w.seg_cap;
    float errmax = 0.f;
    int nre = 0;
    if (Tmax > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (0 < Tr[i]) ax_cp_async16(wst + (i * 4 + prow) * AX_FS_ROW + piece * 8, reinterpret_cast<const void*>(src[i]));
        ax_cp_async_commit();
    }
    for (int t = 0; t <= Tmax; ++t) {
        float* ycur = wy + (t & 1) * (AX_FS_THREADS * AX_F32_YROW);
        float* yprev = wy + ((t & 1) ^ 1) * (AX_FS_THREADS * AX_F32_YROW);
        if (t < Tmax) {
            if (t + 1 < Tmax) {
                const int s = (t + 1) & 1;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (t + 1 < Tr[i]) ax_cp_async16(wst + s * AX_FS_STAGE + (i * 4 + prow) * AX_FS_ROW + piece * 8,
                                                     reinterpret_cast<const void*>(src[i] + (unsigned long long)(t + 1) * 128));
                ax_cp_async_commit();
                ax_cp_async_wait<1>();
            } else ax_cp_async_wait<0>();
            __syncwarp();
            // ---------------- phase 1: fp32 cascade over this thread's 64 samples
            if (t < T) {
                const int4* rp = reinterpret_cast<const int4*>(wst + (t & 1) * AX_FS_STAGE + lane * AX_FS_ROW);
                float* yo = ycur + lane * AX_F32_YROW;
#pragma unroll 1
                for (int v = 0; v < 8; ++v) {
                    const int4 q = rp[v];
                    const int wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
#pragma unroll
                        for (int e2 = 0; e2 < 2; ++e2) {
                            const int xi = e2 ? (wd[h] >> 16) : (int)(short)(wd[h] & 0xFFFF);
                            float tt = fmaf((float)xi, k0, k1);
#pragma unroll
                            for (int s = 0; s < NSEC; ++s) {
                                const float y = tt + z0[s];
                                z0[s] = fmaf(-a1[s], y, fmaf(sg[s], tt, z1[s]));
                                z1[s] = fmaf(-a2[s], y, tt);
                                tt = y;
                            }
                            yo[v * 8 + h * 2 + e2] = tt;
                        }
                    }
                }
            }
        }
        __syncwarp();
        // ---------------- phase 2: warp-cooperative, one thread-row at a time
        for (int r = 0; r < 32; ++r) {
            const int Tq = __shfl_sync(0xffffffffu, T, r);
            if (Tq == 0) continue;
            const int nbq = __shfl_sync(0xffffffffu, nb, r), nstopq = __shfl_sync(0xffffffffu, nstop, r);
            const int sstartq = __shfl_sync(0xffffffffu, sstart, r), sendq = __shfl_sync(0xffffffffu, send, r);
            float* yc = ycur + r * AX_F32_YROW;
            float* yp = yprev + r * AX_F32_YROW;
            // -------- 2a: fp64 re-evaluation of the samples of the current row that are too close to zero
            if (t < Tq) {
                const int base = nbq + 64 * t;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int p = half * 32 + lane, n = base + p;
                    const float y = yc[p];
                    const bool flag = (n < nstopq) && (n >= sstartq) && (n <= sendq) && (fabsf(y) < G);
                    unsigned ball = __ballot_sync(0xffffffffu, flag);
                    while (ball) {
                        const int L = __ffs((int)ball) - 1;
                        ball &= ball - 1;
                        const int nf = base + half * 32 + L;
                        double part = ax_fir_partial(xdrop, nf, c.fir_h, c.fir_len, lane, 32);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                        if (lane == 0) {
                            const double y64 = ax_fir_finish(part, nf, c, st);
                            const float old = yc[half * 32 + L];
                            float fixed = (float)y64;
                            if (fixed == 0.f && y64 != 0.0) fixed = copysignf(1e-37f, (float)(y64 < 0.0 ? -1.0 : 1.0));
                            yc[half * 32 + L] = fixed;
                            errmax = fmaxf(errmax, fabsf((float)(y64 - (double)old)));
                            ++nre;
                        }
                    }
                }
                __syncwarp();
            }
            // -------- 2b: crossings of the previous row and their mark / space windows
            if (t >= 1 && t - 1 < Tq) {
                const int base = nbq + 64 * (t - 1);
                const bool have_cur = t < Tq;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int p = half * 32 + lane, i = base + p;
                    const float y0 = yp[p];
                    const float y1 = (p < 63) ? yp[p + 1] : (have_cur ? yc[0] : 0.f);
                    const bool ok = (i >= sstartq) && (i < sendq) && (i + 1 < nstopq);
                    unsigned ball = __ballot_sync(0xffffffffu, ok && ((y0 < 0.f) != (y1 < 0.f)));
                    while (ball) {
                        const int L = __ffs((int)ball) - 1;
                        ball &= ball - 1;
                        const int pc = half * 32 + L;            // crossing position inside the previous row
                        float sr1 = 0.f, si1 = 0.f, sr2 = 0.f, si2 = 0.f;
                        bool complete = (base + pc + npcm < nstopq);
#pragma unroll
                        for (int rep = 0; rep < 2; ++rep) {
                            const int m = rep * 32 + lane;
                            if (m < npcm) {
                                const int qpos = pc + 1 + m;
                                const float yv = (qpos < 64) ? yp[qpos] : (have_cur ? yc[qpos - 64] : 0.f);
                                const float4 tb = tabf[m];
                                sr1 = fmaf(yv, tb.x, sr1); si1 = fmaf(yv, tb.y, si1);
                                sr2 = fmaf(yv, tb.z, sr2); si2 = fmaf(yv, tb.w, si2);
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            sr1 += __shfl_xor_sync(0xffffffffu, sr1, o); si1 += __shfl_xor_sync(0xffffffffu, si1, o);
                            sr2 += __shfl_xor_sync(0xffffffffu, sr2, o); si2 += __shfl_xor_sync(0xffffffffu, si2, o);
                        }
                        if (lane == 0) {
                            const int cnt = count[warp * 32 + r];
                            if (cnt < w.seg_cap) {
                                const int64_t o = slot0 + (int64_t)(r - lane) * w.seg_cap + cnt;      // slot of thread-row r (lane == 0 here)
                                w.rec_idx[o] = base + pc;
                                w.rec_a1[o] = complete ? (double)sqrtf(sr1 * sr1 + si1 * si1) : ax_nan();
                                w.rec_a2[o] = complete ? (double)sqrtf(sr2 * sr2 + si2 * si2) : ax_nan();
                            }
                            count[warp * 32 + r] = cnt + 1;
                        }
                    }
                }
            }
        }
        __syncwarp();
    }
    __syncwarp();
    if (active) {
        int cnt = count[threadIdx.x];
        if (cnt > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; cnt = w.seg_cap; }
        w.seg_cnt[seg] = cnt;
    } else w.seg_cnt[seg] = 0;
    if (lane == 0 && nre > 0) {
        atomicAdd(&st.n_recheck, nre);
        atomicMax(&st.err32_bits, __float_as_int(errmax));
    }
}

template <int NSEC>
static inline void ax_launch_filter32(const AxWave& w, cudaStream_t stream) {
    const size_t smem = 64 * sizeof(float4) + 2 * AX_FS_STAGE * sizeof(int16_t) +
                        2 * AX_FS_THREADS * AX_F32_YROW * sizeof(float) + AX_FS_THREADS * sizeof(int);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_filter32<NSEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr_set = true; }
    k_filter32<NSEC><<<w.nseg_total / AX_FS_THREADS, AX_FS_THREADS, smem, stream>>>(w);
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

// ------------------------------------------------------------------ frame sync (warp per drop)
// Same greedy scan as ax_frames_item (parse.py:57-89 over AXCTDprocessor.py's per-iteration buffers),
// but one warp walks a drop: the 32 lanes fetch 1024 mask positions per load, so the chain of
// dependent memory latencies is one per 32 frames instead of one per frame.  Control flow is
// warp-uniform; lane 0 records the frame positions.
__global__ void __launch_bounds__(128) k_frames_warp(AxWave w) {
    const int d = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (d >= w.n_drops) return;
    const int lane = threadIdx.x & 31;
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (lane == 0) st.n_frames = 0;
    if (st.status != 0 || st.sm_status < 2 || st.k2 < 0 || st.nedges_total == 0) return;
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int32_t* I = w.edge_idx + dr.edge_base;
    const uint32_t* vw = w.validw + dr.edge_base / 32;
    const int64_t nwords = (st.nbits_total + 31) / 32;
    axctd_frame* fr = w.frame + dr.frame_base;
    const int64_t prof = st.profstartind;
    int64_t cur = 0, wbase = -1;
    uint32_t myword = 0;
    int32_t nf = 0;
    for (int k = st.k2; k < st.n_chunks; ++k) {
        if (lane == 0) { ch[k].frame_begin = nf; ch[k].frame_end = nf; }
        if (ch[k].n_edges <= 0) continue;
        const int64_t NI = ch[k].edge_off + ch[k].n_edges, NB = ch[k].bit_off + ch[k].n_edges - 1;
        if (cur < NI && (int64_t)I[cur] <= prof) {
            const int64_t f = ax_first_gt(I, cur, NI, prof);
            if (f < 0) { if (lane == 0) ax_raise(st, AXCTD_DROP_TRIM_INDEX, k); return; }
            cur = f;
        }
        const int64_t limit = NB - 32;
        int64_t p = cur;
        while (p < limit) {
            const int64_t wi = p >> 5;
            if (wbase < 0 || wi < wbase || wi >= wbase + 32) {
                wbase = wi;
                myword = (wbase + lane < nwords) ? vw[wbase + lane] : 0u;
            }
            uint32_t v = myword;
            const int64_t mine = wbase + lane;
            if (mine < wi) v = 0u; else if (mine == wi) v &= ~((1u << (p & 31)) - 1u);
            const unsigned ball = __ballot_sync(0xffffffffu, v != 0u);
            if (ball == 0u) { p = (wbase + 32) << 5; if (p > limit) p = limit; continue; }
            const int L = __ffs((int)ball) - 1;
            const uint32_t vv = __shfl_sync(0xffffffffu, v, L);
            const int64_t pos = ((wbase + L) << 5) + (__ffs((int)vv) - 1);
            if (pos >= limit) { p = limit; break; }
            if (nf >= dr.frame_cap) { if (lane == 0) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; } return; }
            if (lane == 0) { fr[nf].edge_index = pos; fr[nf].chunk = k; }
            ++nf;
            p = pos + 32;
        }
        if (p > cur) cur = p;                      // AXCTDprocessor.py:618-621
        if (lane == 0) ch[k].frame_end = nf;
    }
    if (lane == 0) st.n_frames = nf;
}

static inline void ax_launch_frames_warp(const AxWave& w, cudaStream_t stream) {
    const int warps_per_block = 4;
    k_frames_warp<<<(w.n_drops + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(w);
}
