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

// ------------------------------------------------------------------ tones
#define AX_TONE_WARPS 8
#define AX_TONE_R 4          // blocks per warp pass (register blocking against the smem table)

__device__ __forceinline__ bool ax_tone_chunk_active(const AxState& st, int k, int phase_b) {
    if (!phase_b) return k < st.n_fixed;
    return st.sm_status >= 1 && k > st.k0 && k < st.n_chunks;
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
        if (k >= dr.chunk_cap || st.status >= AXCTD_DROP_CAPACITY || !ax_tone_chunk_active(st, k, phase_b)) continue;
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
    if (!phase_b) { klo = 0; khi = st.n_fixed; }
    else { if (st.sm_status < 1) return; klo = st.k0 + 1; khi = st.n_chunks; }
    if (khi <= klo) return;
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
