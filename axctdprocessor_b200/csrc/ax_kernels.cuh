// ax_kernels.cuh -- CUDA-only cooperative kernels (sm_100a).
//
//   k_stats_tones     one PCM pass: sum / min / max of the int16 samples (AXCTDprocessor.py:55-56) and the
//                     400 / 7500 / dead-frequency DFT sums of every aligned 256-sample block (:358-364)
//   k_tone_windows    0.1 s window magnitudes from the block sums and the ragged ends
//   k_demod_fused     SOS cascade, zero crossings, mark / space windows (demodulate.py:74-102)
//   k_compact_warp / k_canon_block               dense crossing arrays + walk steps, canonical walk tables
//   k_emit_chunk / k_bits_chunk                  bit edges and bit decisions, one CTA per run() iteration
#pragma once
#include <cuda_runtime.h>
#include "ax_proto.h"

__device__ __forceinline__ void ax_cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void ax_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void ax_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- bulk asynchronous copies (the TMA unit's 1-D form: cp.async.bulk, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ unsigned ax_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ax_mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(ax_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void ax_mbar_arrive(unsigned long long* b) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(ax_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ax_mbar_wait(unsigned long long* b, unsigned parity) {
    const unsigned a = ax_smem_u32(b);
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// one arrival that also announces `bytes` of asynchronous copies to come
__device__ __forceinline__ void ax_mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(ax_smem_u32(b)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes `bytes` on the mbarrier
__device__ __forceinline__ void ax_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(ax_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(ax_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void ax_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per device, so it is tracked per
// (kernel, device): an engine on a second GPU of the same process sets it again.
#include <atomic>
template <auto Kernel>
static inline cudaError_t ax_optin_smem(size_t bytes, int device) {
    static std::atomic<unsigned long long> done{0ull};
    const unsigned long long bit = 1ull << (device & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const cudaError_t r = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (r == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return r;
}

// ------------------------------------------------------------------ stats (exact |x| semantics, fallback)
// np.max(np.abs(int16)) wraps abs(-32768) to -32768 (AXCTDprocessor.py:56).  k_stats_tones tracks min and max,
// which decide max|x| unless a sample equals -32768; only then this kernel rescans the drop.
// grid (AX_WRAP_CTAS, drops): the CTAs of a drop share its slabs; almost always they only look at vmin and leave
#define AX_WRAP_CTAS 16
__global__ void __launch_bounds__(256) k_stats_wrap(AxWave w) {
    const int d = blockIdx.y;
    if (w.st[d].vmin != -32768) return;
    const AxDrop& dr = w.drop[d];
    int mx = -32768;
    for (int64_t j = blockIdx.x; j < dr.nslab; j += gridDim.x) {
        const int64_t a = j * AX_STAT_SLAB;
        int64_t b = a + AX_STAT_SLAB;
        if (b > dr.n_raw) b = dr.n_raw;
        const int16_t* x = w.pcm + dr.pcm_off + a;
        for (int t = threadIdx.x; t < (int)(b - a); t += blockDim.x) {
            const int v = x[t];
            mx = max(mx, (v == -32768) ? -32768 : abs(v));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&w.st[d].ampl, mx);
}
// k_init with a warp per drop: the state record (1.3 KB) is cleared with 16-byte stores by all lanes
__global__ void __launch_bounds__(32) k_init_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    AxState& st = w.st[d];
    static_assert(sizeof(AxState) % 8 == 0, "AxState is cleared in 8-byte words");
    unsigned long long* p = reinterpret_cast<unsigned long long*>(&st);
    for (int i = lane; i < (int)(sizeof(AxState) / 8); i += 32) p[i] = 0ull;
    __syncwarp();
    if (lane == 0) ax_state_defaults(st, w.cfg[w.drop[d].cfg]);
}

// ------------------------------------------------------------------ stats + tone block sums (one PCM pass)
// Sum, min and max of the int16 samples (AXCTDprocessor.py:55-56) and, for every aligned block of AX_TB
// samples, the raw 400 / 7500 / dead-frequency DFT sums (ax_toneblock_item; AXCTDprocessor.py:358-364).
// One lane per tone block; the samples are staged as in k_demod_fused (16-byte cp.async, one 128-byte
// line per lane per stage) and the phasors are constant-bank operands of the DFMAs.
#define AX_ST_THREADS 128
__global__ void __launch_bounds__(AX_ST_THREADS) k_stats_tones(const __grid_constant__ AxWave w, const __grid_constant__ AxToneTab tab, int cfg_id) {
    __shared__ __align__(16) int16_t stage_all[AX_ST_THREADS / 32][2][32 * 72];
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;
    const int64_t nsamp = dr.n_raw;                  // statistics are taken over the recording as uploaded
    const int64_t nblk = (nsamp + AX_TB - 1) / AX_TB;
    if ((int64_t)blockIdx.x * AX_ST_THREADS >= nblk) return;
    if (w.streaming && ((int64_t)blockIdx.x + 1) * AX_ST_THREADS <= w.st[d].tb_done) return;     // summed by an earlier run
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t (*stage)[32 * 72] = stage_all[warp];
    const int64_t jb = (int64_t)blockIdx.x * AX_ST_THREADS + threadIdx.x;
    const bool active = jb < nblk;
    const int64_t n0 = jb * AX_TB;
    const int T = active ? (int)min((int64_t)4, (nsamp - n0 + 63) >> 6) : 0;
    const unsigned long long xrow = (unsigned long long)(w.pcm + dr.pcm_off + n0);
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
            if (t < Tr[i]) ax_cp_async16(&stage[s][(i * 4 + prow) * 72 + piece * 8], reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
        ax_cp_async_commit();
    };
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    long long sum = 0;
    int mx2 = (int)0x80008000, mn2 = 0x7fff7fff;        // packed int16 max / min
    issue(0, 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r + 1 < 4) { issue(r + 1, (r + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
        __syncwarp();
        if (r < T) {
            const int4* rp = reinterpret_cast<const int4*>(&stage[r & 1][lane * 72]);
            const int nvalid = (int)min((int64_t)64, nsamp - (n0 + 64 * r));
            if (nvalid == 64) {
                int s32 = 0;
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int4 q = rp[v];
                    const int wd[4] = {q.x, q.y, q.z, q.w};
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)wd[0], (unsigned)wd[1]);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)wd[2], (unsigned)wd[3]);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)wd[0], (unsigned)wd[1]);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)wd[2], (unsigned)wd[3]);
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        s32 = __dp2a_lo(wd[h], 0x0101, s32);
                        const short2 xs2 = *reinterpret_cast<const short2*>(&wd[h]);      // I2F.F64.S16 on either half: no unpacking
                        const double x0 = (double)xs2.x, x1 = (double)xs2.y;
                        const int m = 64 * r + v * 8 + h * 2;
#pragma unroll
                        for (int q6 = 0; q6 < 6; ++q6) acc[q6] = fma(x0, tab.t[m][q6], acc[q6]);
#pragma unroll
                        for (int q6 = 0; q6 < 6; ++q6) acc[q6] = fma(x1, tab.t[m + 1][q6], acc[q6]);
                    }
                }
                sum += s32;
            } else {                                   // ragged end of the recording: statistics only
                const int16_t* xs = &stage[r & 1][lane * 72];
                for (int i = 0; i < nvalid; ++i) {
                    const int v = xs[i];
                    sum += v;
                    const int pk = (v & 0xFFFF) | (v << 16);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)pk, (unsigned)pk);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)pk, (unsigned)pk);
                }
            }
        }
        __syncwarp();
    }
    if (active && jb < dr.ntb && dr.xf_off < 0) {      // (a decimating drop takes its block sums from the halved signal)
        double* out = w.tb_sum + (dr.tb_base + jb) * 6;
#pragma unroll
        for (int q6 = 0; q6 < 6; ++q6) out[q6] = acc[q6];
    }
    int mx = max((int)(short)(mx2 & 0xFFFF), mx2 >> 16), mn = min((int)(short)(mn2 & 0xFFFF), mn2 >> 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&w.st[d].sum, (unsigned long long)sum);
        atomicMax(&w.st[d].vmax, mx);
        atomicMin(&w.st[d].vmin, mn);
    }
}

// The same pass with the block sums on the FP64 tensor cores.  The sums are a tall-skinny product -- (blocks x 256
// samples) x (256 x 6 phasors) -- and one DMMA m8n8k4 does the work of eight warp-wide DFMAs (eight blocks x four
// samples x eight columns, six of them used) while holding the issue port for a single cycle: a vector DFMA blocks
// it for 2.2 cycles, which is what bounds k_stats_tones (tools/ubench.cu: 16.5 cycles per DMMA with up to ~12 other
// instructions issued underneath).  Lane = tone block for the staging and the statistics as before; for the DMMAs
// the warp's 32 blocks form four groups of eight rows: A[i][c] = sample 4 ks + c of block 8 g + i (lane = 4 i + c),
// B[c][n] = phasor n of that sample (shared-memory table, one 8-byte load per k-step, reused by the four groups),
// C[i][2 (lane & 3) + {0, 1}] accumulates over the 64 k-steps of a block.
#define AX_STM_GROUPS 4
#define AX_STM_SMEM (AX_TB * 8 * sizeof(double) + (AX_ST_THREADS / 32) * 2 * 32 * 72 * sizeof(int16_t))
__device__ __forceinline__ void ax_dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(AX_ST_THREADS) k_stats_tones_mma(const __grid_constant__ AxWave w, const double* __restrict__ tab8, int cfg_id) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];       // AX_STM_SMEM bytes: phasor table, then the staging rows
    double* ptab = reinterpret_cast<double*>(ax_smem_raw);
    int16_t (*stage_all)[2][32 * 72] = reinterpret_cast<int16_t (*)[2][32 * 72]>(ax_smem_raw + AX_TB * 8 * sizeof(double));
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;
    const int64_t nsamp = dr.n_raw;                  // statistics are taken over the recording as uploaded
    const int64_t nblk = (nsamp + AX_TB - 1) / AX_TB;
    if ((int64_t)blockIdx.x * AX_STM_GROUPS * AX_ST_THREADS >= nblk) return;
    if (w.streaming && ((int64_t)blockIdx.x + 1) * AX_STM_GROUPS * AX_ST_THREADS <= w.st[d].tb_done) return;   // summed by an earlier run
    for (int i = threadIdx.x; i < AX_TB * 8; i += AX_ST_THREADS) ptab[i] = tab8[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t (*stage)[32 * 72] = stage_all[warp];
    const int prow = lane >> 3, piece = lane & 7;
    const int arow = lane >> 2, acol = lane & 3;     // A / C row (block within the group), A column (sample within the k-step)
    long long sum = 0;
    int mx2 = (int)0x80008000, mn2 = 0x7fff7fff;        // packed int16 max / min
    // a CTA takes AX_STM_GROUPS consecutive groups of 128 blocks, so that the table load above is paid once per 128 K samples
#pragma unroll 1
    for (int grp = 0; grp < AX_STM_GROUPS; ++grp) {
    const int64_t jg = ((int64_t)blockIdx.x * AX_STM_GROUPS + grp) * AX_ST_THREADS;
    if (jg >= nblk) break;
    const int64_t jb = jg + threadIdx.x;
    const bool active = jb < nblk;
    const int64_t n0 = jb * AX_TB;
    const int T = active ? (int)min((int64_t)4, (nsamp - n0 + 63) >> 6) : 0;
    const unsigned long long xrow = (unsigned long long)(w.pcm + dr.pcm_off + n0);
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
            if (t < Tr[i]) ax_cp_async16(&stage[s][(i * 4 + prow) * 72 + piece * 8], reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
        ax_cp_async_commit();
    };
    double acc[4][2];
#pragma unroll
    for (int g = 0; g < 4; ++g) { acc[g][0] = 0.0; acc[g][1] = 0.0; }
    issue(0, 0);
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        if (r + 1 < 4) { issue(r + 1, (r + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
        __syncwarp();
        // block sums: rows of blocks that do not reach this far hold stale samples; their sums are never stored
        {
            const int16_t* st = stage[r & 1];
            const double* pt = ptab + (64 * r + acol) * 8 + arow;
#pragma unroll 4
            for (int ks = 0; ks < 16; ++ks) {
                const double b = pt[32 * ks];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const double a = (double)st[(8 * g + arow) * 72 + 4 * ks + acol];
                    ax_dmma884(acc[g][0], acc[g][1], a, b);
                }
            }
        }
        if (r < T) {
            const int4* rp = reinterpret_cast<const int4*>(&stage[r & 1][lane * 72]);
            const int nvalid = (int)min((int64_t)64, nsamp - (n0 + 64 * r));
            if (nvalid == 64) {
                int s32 = 0;
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int4 q = rp[v];
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.x, (unsigned)q.y);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.z, (unsigned)q.w);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.x, (unsigned)q.y);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.z, (unsigned)q.w);
                    s32 = __dp2a_lo(q.x, 0x0101, s32); s32 = __dp2a_lo(q.y, 0x0101, s32);
                    s32 = __dp2a_lo(q.z, 0x0101, s32); s32 = __dp2a_lo(q.w, 0x0101, s32);
                }
                sum += s32;
            } else {                                   // ragged end of the recording
                const int16_t* xs = &stage[r & 1][lane * 72];
                for (int i = 0; i < nvalid; ++i) {
                    const int v = xs[i];
                    sum += v;
                    const int pk = (v & 0xFFFF) | (v << 16);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)pk, (unsigned)pk);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)pk, (unsigned)pk);
                }
            }
        }
        __syncwarp();
    }
    if (dr.xf_off < 0 && acol < 3) {                   // (a decimating drop takes its block sums from the halved signal)
        const int64_t jw = jg + warp * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int64_t j = jw + 8 * g + arow;
            if (j < dr.ntb) {
                double* out = w.tb_sum + (dr.tb_base + j) * 6 + 2 * acol;
                out[0] = acc[g][0]; out[1] = acc[g][1];
            }
        }
    }
    }
    int mx = max((int)(short)(mx2 & 0xFFFF), mx2 >> 16), mn = min((int)(short)(mn2 & 0xFFFF), mn2 >> 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&w.st[d].sum, (unsigned long long)sum);
        atomicMax(&w.st[d].vmax, mx);
        atomicMin(&w.st[d].vmin, mn);
    }
}

// The same pass with the block sums as EXACT INTEGER products on the int8 tensor cores (IMMA.16832, 8.4 cycles per
// instruction and scheduler on B200 against 16.5 for a DMMA that does a sixteenth of the multiply-adds:
// tools/ubench_imma.cu).  A sample is two 8-bit slices, x = 256 xh + xl (xh signed, xl unsigned); a phasor is
// quantised to P = round(p * 2^45) and written with six signed base-256 digits, P = sum_j d_j 256^j, d_j in
// [-128, 127] (axctd_config_create).  Every product slice x digit is an m16n8k32 tile product -- 16 tone blocks x 32
// samples x 8 columns (six used) -- accumulated exactly in int32 over the block's 256 samples (|sum| < 2^24); slices
// of equal weight share an accumulator (xh d_(e-1) and xl d_e both weigh 256^e), and the block sum is
// sum_e acc_e 256^e * 2^-45, formed in double at the end.  The result is the exact sum of x[n] P[n]: it differs from
// the double-precision sum only by the quantisation of the phasors, <= 2^-46 each, i.e. <= 1.2e-7 on a block sum of
// full-scale samples (relative 1e-13; the FP64 sum itself rounds at 1e-10 absolute on such a block).  Twelve IMMAs per
// tile and k-step = 0.023 per sample: 0.7 ms of tensor time for the 128-drop batch, below the 1.3 ms its PCM takes to
// stream from HBM -- the pass becomes bandwidth-bound.
// Fragments (PTX m16n8k32, lane = 4 g + t): A rows g and g + 8 of the tile, bytes k = 4 t .. 4 t + 3 and 16 + 4 t ..
// (one 8-byte shared-memory load of four staged int16 samples per register, split into low and high bytes with two
// PRMT); B column g, same k (digit table in shared memory, laid out per lane); C rows g, g + 8, columns 2 t, 2 t + 1.
#define AX_STI_DIGITS 6
#define AX_STI_SHIFT 45
#define AX_STI_STAGES 2                                 // staging rows per warp: one in flight while one is consumed (three were measured: no gain)
#define AX_STI_GROUPS 8                                 // groups of 128 tone blocks per CTA (the digit table is loaded once per CTA)
#define AX_STI_TAB_WORDS ((AX_TB / 32) * AX_STI_DIGITS * 32 * 2)
#define AX_STI_SMEM (AX_STI_TAB_WORDS * sizeof(uint32_t) + (AX_ST_THREADS / 32) * AX_STI_STAGES * 32 * 72 * sizeof(int16_t))
__device__ __forceinline__ void ax_imma_u8s8(int (&c)[4], const unsigned (&a)[4], const uint2 b) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
__device__ __forceinline__ void ax_imma_s8s8(int (&c)[4], const unsigned (&a)[4], const uint2 b) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}
// A CTA takes AX_STI_GROUPS consecutive groups of 128 tone blocks (a warp: 32 blocks of each group, lane = block for the
// staging and the statistics); the 4 x AX_STI_GROUPS row steps of a warp form ONE software pipeline -- the rows of the
// next group are already in flight while the last rows of a group are consumed -- with AX_STI_STAGES - 1 rows ahead.
__global__ void __launch_bounds__(AX_ST_THREADS, 4) k_stats_tones_imma(const __grid_constant__ AxWave w, const uint32_t* __restrict__ tabi, int cfg_id) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];       // AX_STI_SMEM bytes: digit table, then the staging rows
    const uint2* btab = reinterpret_cast<const uint2*>(ax_smem_raw);   // [k-step of the block][digit][lane]
    int16_t (*stage_all)[AX_STI_STAGES][32 * 72] = reinterpret_cast<int16_t (*)[AX_STI_STAGES][32 * 72]>(ax_smem_raw + AX_STI_TAB_WORDS * sizeof(uint32_t));
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;
    const int64_t nsamp = dr.n_raw;                  // statistics are taken over the recording as uploaded
    const int64_t nblk = (nsamp + AX_TB - 1) / AX_TB;
    if ((int64_t)blockIdx.x * AX_STI_GROUPS * AX_ST_THREADS >= nblk) return;
    if (w.streaming && ((int64_t)blockIdx.x + 1) * AX_STI_GROUPS * AX_ST_THREADS <= w.st[d].tb_done) return;   // summed by an earlier run
    // digit table: asynchronous 16-byte copies that land together with the first staged row (first commit group)
    for (int i = threadIdx.x; i < AX_STI_TAB_WORDS / 4; i += AX_ST_THREADS) ax_cp_async16(ax_smem_raw + 16 * i, tabi + 4 * i);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t (*stage)[32 * 72] = stage_all[warp];
    const int prow = lane >> 3, piece = lane & 7;
    const int fg = lane >> 2, ft = lane & 3;         // fragment row / column group
    long long sum = 0;
    int mx2 = (int)0x80008000, mn2 = 0x7fff7fff;        // packed int16 max / min
    const int16_t* xdrop = w.pcm + dr.pcm_off;
    const int64_t jw0 = (int64_t)blockIdx.x * AX_STI_GROUPS * AX_ST_THREADS + warp * 32;      // first block of this warp in group 0
    // row step s = 4 grp + r: row r (64 samples) of the warp's 32 blocks of group grp; a lane copies the 16-byte piece
    // `piece` of the rows of blocks 4 i + prow, i < 8 (blocks are 512 bytes apart: no address exchange needed)
    auto issue = [&](int s, int slot) {
        const int64_t jw = jw0 + (int64_t)(s >> 2) * AX_ST_THREADS;
        const int r = s & 3;
        const int64_t rem = nsamp - (jw + prow) * AX_TB - 64 * r;        // samples left from this row's start, block prow
        const int16_t* src = xdrop + (jw + prow) * AX_TB + 64 * r + piece * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (rem - (int64_t)i * 4 * AX_TB > 0) ax_cp_async16(&stage[slot][(i * 4 + prow) * 72 + piece * 8], src + (int64_t)i * 4 * AX_TB);
        ax_cp_async_commit();
    };
    int nsteps = 0;
#pragma unroll
    for (int grp = 0; grp < AX_STI_GROUPS; ++grp)
        if (((int64_t)blockIdx.x * AX_STI_GROUPS + grp) * AX_ST_THREADS < nblk) nsteps = 4 * (grp + 1);
    int acc[2][AX_STI_DIGITS + 1][4];                // [tile of 16 blocks][weight 256^e][C fragment]
#pragma unroll
    for (int pre = 0; pre < AX_STI_STAGES - 1; ++pre) { if (pre < nsteps) issue(pre, pre); else ax_cp_async_commit(); }
    int slot = 0, slot_next = AX_STI_STAGES - 1;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        const int r = s & 3;
        const int64_t jw = jw0 + (int64_t)(s >> 2) * AX_ST_THREADS;
        if (r == 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int e = 0; e <= AX_STI_DIGITS; ++e) { acc[mt][e][0] = 0; acc[mt][e][1] = 0; acc[mt][e][2] = 0; acc[mt][e][3] = 0; }
        }
        if (s + AX_STI_STAGES - 1 < nsteps) issue(s + AX_STI_STAGES - 1, slot_next); else ax_cp_async_commit();
        ax_cp_async_wait<AX_STI_STAGES - 1>();
        if (s == 0) __syncthreads(); else __syncwarp();      // (step 0: the digit table, copied by all warps, has landed too)
        // block sums: rows of blocks that do not reach this far hold stale samples; their sums are never stored
        {
            const int16_t* st = stage[slot];
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                const uint2* bt = btab + ((2 * r + k2) * AX_STI_DIGITS) * 32 + lane;
                uint2 bd[AX_STI_DIGITS];
#pragma unroll
                for (int j = 0; j < AX_STI_DIGITS; ++j) bd[j] = bt[32 * j];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int16_t* p0 = st + (16 * mt + fg) * 72 + 32 * k2 + 4 * ft;
                    const uint2 w00 = *reinterpret_cast<const uint2*>(p0), w01 = *reinterpret_cast<const uint2*>(p0 + 16);
                    const uint2 w10 = *reinterpret_cast<const uint2*>(p0 + 8 * 72), w11 = *reinterpret_cast<const uint2*>(p0 + 8 * 72 + 16);
                    unsigned al[4], ah[4];
                    al[0] = __byte_perm(w00.x, w00.y, 0x6420); ah[0] = __byte_perm(w00.x, w00.y, 0x7531);
                    al[1] = __byte_perm(w10.x, w10.y, 0x6420); ah[1] = __byte_perm(w10.x, w10.y, 0x7531);
                    al[2] = __byte_perm(w01.x, w01.y, 0x6420); ah[2] = __byte_perm(w01.x, w01.y, 0x7531);
                    al[3] = __byte_perm(w11.x, w11.y, 0x6420); ah[3] = __byte_perm(w11.x, w11.y, 0x7531);
                    // (the two products that meet in one accumulator are issued six instructions apart)
#pragma unroll
                    for (int j = 0; j < AX_STI_DIGITS; ++j) ax_imma_u8s8(acc[mt][j], al, bd[j]);
#pragma unroll
                    for (int j = 0; j < AX_STI_DIGITS; ++j) ax_imma_s8s8(acc[mt][j + 1], ah, bd[j]);
                }
            }
        }
        {   // statistics of this lane's own block
            const int64_t n0 = (jw + lane) * AX_TB + 64 * r;
            const int64_t left = nsamp - n0;
            if (left > 0) {
                const int4* rp = reinterpret_cast<const int4*>(&stage[slot][lane * 72]);
                if (left >= 64) {
                    int s32 = 0;
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int4 q = rp[v];
                        mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.x, (unsigned)q.y);
                        mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.z, (unsigned)q.w);
                        mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.x, (unsigned)q.y);
                        mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.z, (unsigned)q.w);
                        s32 = __dp2a_lo(q.x, 0x0101, s32); s32 = __dp2a_lo(q.y, 0x0101, s32);
                        s32 = __dp2a_lo(q.z, 0x0101, s32); s32 = __dp2a_lo(q.w, 0x0101, s32);
                    }
                    sum += s32;
                } else {                                   // ragged end of the recording
                    const int16_t* xs = &stage[slot][lane * 72];
                    for (int i = 0; i < (int)left; ++i) {
                        const int v = xs[i];
                        sum += v;
                        const int pk = (v & 0xFFFF) | (v << 16);
                        mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)pk, (unsigned)pk);
                        mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)pk, (unsigned)pk);
                    }
                }
            }
        }
        __syncwarp();
        if (r == 3 && dr.xf_off < 0 && ft < 3) {       // (a decimating drop takes its block sums from the halved signal)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {          // rows fg and fg + 8 of the tile
                    const int64_t j = jw + 16 * mt + fg + 8 * h;
                    if (j < dr.ntb) {
                        double v0 = 0.0, v1 = 0.0;
#pragma unroll
                        for (int e = AX_STI_DIGITS; e >= 0; --e) {      // exact conversions, weights 2^(8 e - shift)
                            const double wgt = __longlong_as_double((long long)(1023 + 8 * e - AX_STI_SHIFT) << 52);
                            v0 = fma((double)acc[mt][e][2 * h], wgt, v0);
                            v1 = fma((double)acc[mt][e][2 * h + 1], wgt, v1);
                        }
                        double* out = w.tb_sum + (dr.tb_base + j) * 6 + 2 * ft;
                        out[0] = v0; out[1] = v1;
                    }
                }
        }
        slot_next = slot;
        slot = (slot + 1 == AX_STI_STAGES) ? 0 : slot + 1;
    }
    int mx = max((int)(short)(mx2 & 0xFFFF), mx2 >> 16), mn = min((int)(short)(mn2 & 0xFFFF), mn2 >> 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&w.st[d].sum, (unsigned long long)sum);
        atomicMax(&w.st[d].vmax, mx);
        atomicMin(&w.st[d].vmin, mn);
    }
}

// The same pass with the block sums split between the FP64 tensor cores and the vector FP64 pipe.
template <int KD>
__global__ void __launch_bounds__(AX_ST_THREADS) k_stats_tones_hyb(const __grid_constant__ AxWave w, const double* __restrict__ tab8, const __grid_constant__ AxToneTab tab, int cfg_id) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];       // AX_STM_SMEM bytes: phasor table, then the staging rows
    double* ptab = reinterpret_cast<double*>(ax_smem_raw);
    int16_t (*stage_all)[2][32 * 72] = reinterpret_cast<int16_t (*)[2][32 * 72]>(ax_smem_raw + AX_TB * 8 * sizeof(double));
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    if (dr.cfg != cfg_id) return;
    const int64_t nsamp = dr.n_raw;                  // statistics are taken over the recording as uploaded
    const int64_t nblk = (nsamp + AX_TB - 1) / AX_TB;
    if ((int64_t)blockIdx.x * AX_STM_GROUPS * AX_ST_THREADS >= nblk) return;
    if (w.streaming && ((int64_t)blockIdx.x + 1) * AX_STM_GROUPS * AX_ST_THREADS <= w.st[d].tb_done) return;   // summed by an earlier run
    for (int i = threadIdx.x; i < AX_TB * 8; i += AX_ST_THREADS) ptab[i] = tab8[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t (*stage)[32 * 72] = stage_all[warp];
    const int prow = lane >> 3, piece = lane & 7;
    const int arow = lane >> 2, acol = lane & 3;     // A / C row (block within the group), A column (sample within the k-step)
    long long sum = 0;
    int mx2 = (int)0x80008000, mn2 = 0x7fff7fff;        // packed int16 max / min
    // a CTA takes AX_STM_GROUPS consecutive groups of 128 blocks, so that the table load above is paid once per 128 K samples
#pragma unroll 1
    for (int grp = 0; grp < AX_STM_GROUPS; ++grp) {
    const int64_t jg = ((int64_t)blockIdx.x * AX_STM_GROUPS + grp) * AX_ST_THREADS;
    if (jg >= nblk) break;
    const int64_t jb = jg + threadIdx.x;
    const bool active = jb < nblk;
    const int64_t n0 = jb * AX_TB;
    const int T = active ? (int)min((int64_t)4, (nsamp - n0 + 63) >> 6) : 0;
    const unsigned long long xrow = (unsigned long long)(w.pcm + dr.pcm_off + n0);
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
            if (t < Tr[i]) ax_cp_async16(&stage[s][(i * 4 + prow) * 72 + piece * 8], reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
        ax_cp_async_commit();
    };
    double acc[4][2];
#pragma unroll
    for (int g = 0; g < 4; ++g) { acc[g][0] = 0.0; acc[g][1] = 0.0; }
    double vac[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};      // this lane's block, samples 4 KD .. 63 of every row
    issue(0, 0);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (r + 1 < 4) { issue(r + 1, (r + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
        __syncwarp();
        // block sums: rows of blocks that do not reach this far hold stale samples; their sums are never stored.
        // The first KD k-steps (4 KD samples) of the row go to the tensor cores, the rest to the vector FP64 pipe
        // (lane = block, phasors as constant operands): the DMMA pipe was 80 % busy and the issue port 35 %, so the
        // two run side by side.
        {
            const int16_t* st = stage[r & 1];
            const double* pt = ptab + (64 * r + acol) * 8 + arow;
#pragma unroll 2
            for (int ks = 0; ks < KD; ++ks) {
                const double b = pt[32 * ks];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const double a = (double)st[(8 * g + arow) * 72 + 4 * ks + acol];
                    ax_dmma884(acc[g][0], acc[g][1], a, b);
                }
            }
            const int* rw = reinterpret_cast<const int*>(&st[lane * 72]);
#pragma unroll
            for (int m = 4 * KD; m < 64; m += 2) {
                const int wd = rw[m >> 1];
                const short2 xs2 = *reinterpret_cast<const short2*>(&wd);
                const double x0 = (double)xs2.x, x1 = (double)xs2.y;
#pragma unroll
                for (int q6 = 0; q6 < 6; ++q6) vac[q6] = fma(x0, tab.t[64 * r + m][q6], vac[q6]);
#pragma unroll
                for (int q6 = 0; q6 < 6; ++q6) vac[q6] = fma(x1, tab.t[64 * r + m + 1][q6], vac[q6]);
            }
        }
        if (r < T) {
            const int4* rp = reinterpret_cast<const int4*>(&stage[r & 1][lane * 72]);
            const int nvalid = (int)min((int64_t)64, nsamp - (n0 + 64 * r));
            if (nvalid == 64) {
                int s32 = 0;
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int4 q = rp[v];
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.x, (unsigned)q.y);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)q.z, (unsigned)q.w);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.x, (unsigned)q.y);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)q.z, (unsigned)q.w);
                    s32 = __dp2a_lo(q.x, 0x0101, s32); s32 = __dp2a_lo(q.y, 0x0101, s32);
                    s32 = __dp2a_lo(q.z, 0x0101, s32); s32 = __dp2a_lo(q.w, 0x0101, s32);
                }
                sum += s32;
            } else {                                   // ragged end of the recording
                const int16_t* xs = &stage[r & 1][lane * 72];
                for (int i = 0; i < nvalid; ++i) {
                    const int v = xs[i];
                    sum += v;
                    const int pk = (v & 0xFFFF) | (v << 16);
                    mx2 = (int)__vimax3_s16x2((unsigned)mx2, (unsigned)pk, (unsigned)pk);
                    mn2 = (int)__vimin3_s16x2((unsigned)mn2, (unsigned)pk, (unsigned)pk);
                }
            }
        }
        __syncwarp();
    }
    // the vector-pipe part of every block joins the tensor-core part through shared memory (the staging rows are free now)
    double* vx = reinterpret_cast<double*>(stage[0]);            // [32 blocks][6]
    __syncwarp();
#pragma unroll
    for (int q6 = 0; q6 < 6; ++q6) vx[lane * 6 + q6] = vac[q6];
    __syncwarp();
    if (dr.xf_off < 0 && acol < 3) {                   // (a decimating drop takes its block sums from the halved signal)
        const int64_t jw = jg + warp * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int64_t j = jw + 8 * g + arow;
            if (j < dr.ntb) {
                double* out = w.tb_sum + (dr.tb_base + j) * 6 + 2 * acol;
                out[0] = acc[g][0] + vx[(8 * g + arow) * 6 + 2 * acol]; out[1] = acc[g][1] + vx[(8 * g + arow) * 6 + 2 * acol + 1];
            }
        }
    }
    __syncwarp();
    }
    int mx = max((int)(short)(mx2 & 0xFFFF), mx2 >> 16), mn = min((int)(short)(mn2 & 0xFFFF), mn2 >> 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (lane == 0) {
        atomicAdd((unsigned long long*)&w.st[d].sum, (unsigned long long)sum);
        atomicMax(&w.st[d].vmax, mx);
        atomicMin(&w.st[d].vmin, mn);
    }
}

// ------------------------------------------------------------------ tone window magnitudes
// Per-drop power-sample range [lo, hi) of a tone launch (the range test of ax_tone_slot_active, once per drop).
__global__ void k_tone_range(AxWave w, int phase_b) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= w.n_drops) return;
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    int32_t lo = 0, hi = 0;
    if (st.status < AXCTD_DROP_CAPACITY && ax_tone_blocked_ok(w.cfg[dr.cfg])) {
        const AxChunk* ch = w.chunk + dr.chunk_base;
        if (!phase_b) {
            const int ka = w.pa_lo, kb = w.pa_hi < st.n_fixed ? w.pa_hi : st.n_fixed;
            if (st.searching && kb > ka) { lo = ch[ka].pw_off; hi = ch[kb - 1].pw_off + ch[kb - 1].np; }
        } else if (st.sm_status >= 1 && st.n_chunks > st.k0 + 1) {
            lo = ch[st.k0].pw_off + ch[st.k0].np; hi = ch[st.n_chunks - 1].pw_off + ch[st.n_chunks - 1].np;
            if (w.streaming && st.next_sm_chunk > st.k0 + 1) lo = st.next_sm_chunk < st.n_chunks ? ch[st.next_sm_chunk].pw_off : hi;
        }
    }
    w.tone_rng[2 * d] = lo; w.tone_rng[2 * d + 1] = hi;
}

__device__ __forceinline__ void ax_warp_sum6(double* a, int lane) {
    // after this call lane 0 holds the sums of components 0..2 in a[0..2], lane 16 those of 3..5 in a[0..2]
    const bool hi = lane >= 16;
    double keep[3], send[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) { keep[q] = hi ? a[3 + q] : a[q]; send[q] = hi ? a[q] : a[3 + q]; }
#pragma unroll
    for (int q = 0; q < 3; ++q) keep[q] += __shfl_xor_sync(0xffffffffu, send[q], 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 3; ++q) keep[q] += __shfl_xor_sync(0xffffffffu, keep[q], o);
#pragma unroll
    for (int q = 0; q < 3; ++q) a[q] = keep[q];
}

// One warp per power sample: ragged ends + block sums (ax_tonewin_partial); the six sums go to tone_acc and
// k_tone_mag turns them into magnitudes with one thread per power sample.
// grid = (warps over the per-drop power-sample range [i_lo, i_hi), drop)
__global__ void __launch_bounds__(256, 4) k_tone_windows(AxWave w, int i_lo, int i_hi) {
    const int d = blockIdx.y;
    const int32_t i = i_lo + (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= i_hi || i < w.tone_rng[2 * d] || i >= w.tone_rng[2 * d + 1]) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    const int lane = threadIdx.x & 31;
    const int64_t slot = (int64_t)dr.pw_base + i;
    const int64_t cstart = w.pw_ind[slot];
    double a[6];
    // int16 source and ragged ends of at most AX_TB samples each (every window except the last ones of a recording):
    // all ragged samples of the lane are fetched before the first use -- they come from HBM (the recording is far
    // larger than L2), and fetched one loop iteration at a time their latency was what bound this kernel.  The
    // sums themselves are ax_tonewin_partial's, term by term.
    const int np = c.n_power;
    int64_t j0 = (cstart + AX_TB - 1) / AX_TB, j1 = (cstart + np) / AX_TB;
    if (j1 > dr.ntb) j1 = dr.ntb;
    if (j1 < j0) j1 = j0;
    const int head_n = (int)(j0 * AX_TB - cstart), tail_off = (int)(j1 * AX_TB - cstart);
    if (dr.xf_off >= 0 || head_n > AX_TB || np - tail_off > AX_TB) {
        ax_tonewin_partial(w, dr, c, cstart, lane, 32, a);
    } else {
        const int16_t* xs = w.pcm + dr.pcm_off + cstart;
        int xh[AX_TB / 32], xt[AX_TB / 32];
#pragma unroll
        for (int k = 0; k < AX_TB / 32; ++k) {
            const int m = lane + 32 * k;
            xh[k] = (m < head_n) ? (int)xs[m] : 0;
            xt[k] = (tail_off + m < np) ? (int)xs[tail_off + m] : 0;
        }
        // phasors from the interleaved table (three 16-byte loads per sample instead of six 8-byte ones)
        const double2* tc = reinterpret_cast<const double2*>(c.tone_cs);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0;
#pragma unroll
        for (int k = 0; k < AX_TB / 32; ++k) {
            const int m = lane + 32 * k;
            if (m < head_n) {
                const double xd = (double)xh[k];
                const double2 p0 = tc[3 * m], p1 = tc[3 * m + 1], p2 = tc[3 * m + 2];
                a0 = fma(xd, p0.x, a0); a1 = fma(xd, p0.y, a1); a2 = fma(xd, p1.x, a2);
                a3 = fma(xd, p1.y, a3); a4 = fma(xd, p2.x, a4); a5 = fma(xd, p2.y, a5);
            }
        }
#pragma unroll
        for (int k = 0; k < AX_TB / 32; ++k) {
            const int m = tail_off + lane + 32 * k;
            if (m < np) {
                const double xd = (double)xt[k];
                const double2 p0 = tc[3 * m], p1 = tc[3 * m + 1], p2 = tc[3 * m + 2];
                a0 = fma(xd, p0.x, a0); a1 = fma(xd, p0.y, a1); a2 = fma(xd, p1.x, a2);
                a3 = fma(xd, p1.y, a3); a4 = fma(xd, p2.x, a4); a5 = fma(xd, p2.y, a5);
            }
        }
        a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3; a[4] = a4; a[5] = a5;
        const int nblk = (int)(j1 - j0);
        const double* B0 = w.tb_sum + (dr.tb_base + j0) * 6;
        double eh[6];
        { const double2 e0 = tc[3 * head_n], e1 = tc[3 * head_n + 1], e2 = tc[3 * head_n + 2];
          eh[0] = e0.x; eh[1] = e0.y; eh[2] = e1.x; eh[3] = e1.y; eh[4] = e2.x; eh[5] = e2.y; }
        for (int jj = lane; jj < nblk; jj += 32) {
            const double* B = B0 + 6 * jj;
            const double* R = c.tone_rot[jj];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const double cr = fma(eh[2 * f], R[2 * f], -(eh[2 * f + 1] * R[2 * f + 1]));
                const double sn = fma(eh[2 * f], R[2 * f + 1], eh[2 * f + 1] * R[2 * f]);
                const double br = B[2 * f], bi = B[2 * f + 1];
                a[2 * f] = fma(br, cr, fma(-bi, sn, a[2 * f]));
                a[2 * f + 1] = fma(br, sn, fma(bi, cr, a[2 * f + 1]));
            }
        }
    }
    ax_warp_sum6(a, lane);
    if ((lane & 15) == 0) { double* o = w.tone_acc + slot * 6 + (lane ? 3 : 0); o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
}
// The same sums with the ragged ends on the FP64 tensor cores: a warp takes eight consecutive power samples.  Their
// ragged heads all use the phasors of window offsets 0 .. 255 and their ragged tails the same phasors relative to the
// tail's first sample (rotated once by e^{j theta tail_off} afterwards, like the block sums), so both are products
// (8 windows x 256 samples) x (256 x 6 phasors): A[i][c] = sample 4 ks + c of window i's head (tail), zero past its end,
// B = the table of k_stats_tones_mma, one DMMA each per k-step, no k-steps past the longest end of the eight.  The
// result lands with lane (i, f) holding (re, im) of frequency f of window i; that lane adds the window's full blocks
// and stores -- no warp reduction, one 2-byte load per ragged sample and no per-sample phasor loads.  Windows with a
// ragged end above AX_TB samples (last windows of a recording) or a double-precision source take the generic sums.
__global__ void __launch_bounds__(128, 8) k_tone_windows_mma(AxWave w, int i_lo, int i_hi) {
    const int d = blockIdx.y;
    const int lo = max(i_lo, w.tone_rng[2 * d]), hi = min(i_hi, w.tone_rng[2 * d + 1]);
    const int lane = threadIdx.x & 31;
    const int i0 = i_lo + 8 * (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    if (i0 >= hi || i0 + 8 <= lo) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    const int arow = lane >> 2, acol = lane & 3;
    const int np = c.n_power;
    const int iw = i0 + arow;
    const bool valid = iw >= lo && iw < hi;
    const int64_t slot = (int64_t)dr.pw_base + iw;
    const int64_t cstart = valid ? w.pw_ind[slot] : 0;
    int64_t j0 = (cstart + AX_TB - 1) / AX_TB, j1 = (cstart + np) / AX_TB;       // full blocks j0 .. j1-1
    if (j1 > dr.ntb) j1 = dr.ntb;
    if (j1 < j0) j1 = j0;
    const int head_n = (int)(j0 * AX_TB - cstart), tail_off = (int)(j1 * AX_TB - cstart);
    const bool regular = valid && dr.xf_off < 0 && head_n <= AX_TB && np - tail_off <= AX_TB && tail_off <= np;
    const int hn = regular ? head_n : 0, tn = regular ? np - tail_off : 0;
    // A ragged end longer than half a block is taken as the block that contains it minus its complement (option
    // tone_complement): the head as e^{-j theta nc} (B_(j0-1) - C), C = the first nc = AX_TB - hn samples of block j0-1
    // against the table; the tail as B_j1 - e^{j theta (AX_TB-1)} conj(R), R = the last nc = AX_TB - tn samples of block
    // j1 read backwards against the table.  No end then needs more than AX_TB / 2 samples: half the k-steps, half the
    // 2-byte sample loads.
    const bool hcomp = w.tone_complement && regular && hn > AX_TB / 2 && j0 >= 1;
    const bool tcomp = w.tone_complement && regular && tn > AX_TB / 2 && j1 < dr.ntb;
    const int hm = hcomp ? AX_TB - hn : hn, tm = tcomp ? AX_TB - tn : tn;
    int kmax = max(hm, tm);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    const int ks_end = (kmax + 3) >> 2;
    const int16_t* xbase = w.pcm + dr.pcm_off;
    const int16_t* xh = hcomp ? xbase + (j0 - 1) * AX_TB : xbase + cstart;           // forwards from here
    const int16_t* xt = tcomp ? xbase + j1 * AX_TB + (AX_TB - 1) : xbase + cstart + tail_off;     // complement: backwards from the block's last sample
    const int tstep = tcomp ? -1 : 1;
    const double* bt = c.tone_tab8 + acol * 8 + arow;
    double h0 = 0.0, h1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll 8
    for (int ks = 0; ks < ks_end; ++ks) {
        const int m = 4 * ks + acol;
        const double b = bt[32 * ks];
        const double ah = (m < hm) ? (double)xh[m] : 0.0;
        const double at = (m < tm) ? (double)xt[tstep * m] : 0.0;
        ax_dmma884(h0, h1, ah, b);
        ax_dmma884(t0, t1, at, b);
    }
    if (regular && acol < 3) {
        const int f = acol;
        const double* tcs = c.tone_cs;
        double are, aim;
        {
            if (hcomp) {
                const double* Bp = w.tb_sum + (dr.tb_base + j0 - 1) * 6 + 2 * f;
                const double dr_ = Bp[0] - h0, di_ = Bp[1] - h1;
                const double cc = tcs[6 * (int64_t)hm + 2 * f], ss = tcs[6 * (int64_t)hm + 2 * f + 1];
                h0 = fma(dr_, cc, di_ * ss);               // (dr + j di) (cc - j ss)
                h1 = fma(di_, cc, -(dr_ * ss));
            }
            if (tcomp) {
                const double* Bp = w.tb_sum + (dr.tb_base + j1) * 6 + 2 * f;
                const double c5 = tcs[6 * (int64_t)(AX_TB - 1) + 2 * f], s5 = tcs[6 * (int64_t)(AX_TB - 1) + 2 * f + 1];
                const double er = fma(t0, c5, t1 * s5), ei = fma(t0, s5, -(t1 * c5));      // (t0 - j t1) (c5 + j s5)
                t0 = Bp[0] - er; t1 = Bp[1] - ei;
            }
            const double ec = tail_off < np ? tcs[6 * (int64_t)tail_off + 2 * f] : 0.0, es = tail_off < np ? tcs[6 * (int64_t)tail_off + 2 * f + 1] : 0.0;
            are = h0 + fma(ec, t0, -(es * t1));
            aim = h1 + fma(ec, t1, es * t0);
        }
        // full blocks: sum_j e^{j theta (head_n + AX_TB j)} B_j = e^{j theta head_n} * sum_j e^{j theta AX_TB j} B_j
        const double ehc = tcs[6 * (int64_t)head_n + 2 * f], ehs = tcs[6 * (int64_t)head_n + 2 * f + 1];
        const int nblk = (int)(j1 - j0);
        const double* B0 = w.tb_sum + (dr.tb_base + j0) * 6 + 2 * f;
        double sr = 0.0, si = 0.0;
#pragma unroll 4
        for (int jj = 0; jj < nblk; ++jj) {
            const double rc = c.tone_rot[jj][2 * f], rs = c.tone_rot[jj][2 * f + 1];
            const double br = B0[6 * jj], bi = B0[6 * jj + 1];
            sr = fma(br, rc, fma(-bi, rs, sr));
            si = fma(br, rs, fma(bi, rc, si));
        }
        are += fma(ehc, sr, -(ehs * si));
        aim += fma(ehc, si, ehs * sr);
        double* o = w.tone_acc + slot * 6 + 2 * f;
        o[0] = are; o[1] = aim;
    }
    // the windows this form does not cover, one after the other with the whole warp
    unsigned irr = __ballot_sync(0xffffffffu, valid && !regular && acol == 0);
    while (irr) {
        const int L = __ffs((int)irr) - 1;
        irr &= irr - 1;
        const int64_t cs = __shfl_sync(0xffffffffu, cstart, L);
        const int64_t sl = (int64_t)dr.pw_base + i0 + (L >> 2);
        double a[6];
        ax_tonewin_partial(w, dr, c, cs, lane, 32, a);
        ax_warp_sum6(a, lane);
        if ((lane & 15) == 0) { double* o = w.tone_acc + sl * 6 + (lane ? 3 : 0); o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
    }
}
__global__ void __launch_bounds__(128) k_tone_mag(AxWave w, int i_lo, int i_hi) {
    const int d = blockIdx.y;
    const int32_t i = i_lo + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= i_hi || i < w.tone_rng[2 * d] || i >= w.tone_rng[2 * d + 1]) return;
    const AxDrop& dr = w.drop[d];
    const int64_t slot = (int64_t)dr.pw_base + i;
    ax_tonewin_finish(w, w.cfg[dr.cfg], w.st[d], slot, w.tone_acc + slot * 6);
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
//   phase 2  every lane takes the crossings of its previous row one after the other and sums their 4*NPCM-FMA
//            fp32 windows (ax_window32, phasors broadcast from shared memory) from its own ring; the warp
//            iterates as often as its busiest lane has crossings (dealing the crossings evenly across the
//            lanes was measured slower: the compaction cost more than the idle lanes).
// y never leaves the SM; the only global traffic is the int16 stream in and (index, |S1|, |S2|) per
// crossing out.
#define AX_FD_WARPS 4
#define AX_FD_THREADS (AX_FD_WARPS * 32)
#define AX_FD_ROW 72                                   // int16 per staged row: 64 samples + 8 pad (144-byte stride)
#define AX_FD_STAGE (32 * AX_FD_ROW)                   // int16 per warp per stage
#define AX_FD_RINGQ 32                                 // quads (4 samples) per lane ring: two rows

struct AxFdWarp {
    AxF4 tab[AX_WIN_TAPS];                             // phasors of the bit windows (per warp copy: no CTA barrier needed)
    int16_t stage[2][AX_FD_STAGE];
    unsigned long long mbar[2];                        // BULK staging: one mbarrier per stage buffer
};
// Shared memory of a CTA: the four y rings first ([warp][quad][lane] float4, 16 KB per warp, so that the byte offset
// of a ring quad is ((quad * 512 + lane * 16) & 0x3ff0) | (warp << 14): one add and one LOP3 per load), then the
// per-warp phasor tables and staging rows.
struct AxFdSmem {
    float4 yring[AX_FD_WARPS][AX_FD_RINGQ * 32];       // every lane reads and writes only its own 16-byte column (no bank conflicts)
    AxFdWarp wp[AX_FD_WARPS];
};
static_assert(AX_FD_RINGQ * 32 * sizeof(float4) == 16384, "ring offsets assume 16 KB per warp");


// HEAD = false: lane = segment of the continuous pass.  HEAD = true: lane = run() iteration, filtered from
// zero state at the chunk start over its first head + npcm + 2 samples (ax_headfilt_item); the lane's
// stream starts at the chunk start rounded down to 8 samples (16-byte staging) and the samples before the
// chunk start enter the cascade as zeros, which leaves its state at zero.
// FAST (low-pass, three sections with numerator (1 + z^-1)^2 each, continuous pass only): from the second row of a
// lane on, the cascade runs "numerators first": u[n] = sum_j C(6,j) x[n-j] is formed exactly in integers (four
// IDP.2A per sample on the packed int16 words, which also replaces the unpacking), tt = k0 u + 64 k1, and the three
// sections are all-pole, y_s[n] = in + a1 y_s[n-1] + a2 y_s[n-2] (two DFMA each): 7 FP64-pipe operations per sample
// instead of 13 -- every one of them costs 2.2 issue cycles on B200 (tools/ubench.cu) -- and one dependent DFMA per
// step and section instead of two.  The result differs from the reference's operation order by < 1e-14 of full
// scale (measured 7.5e-15 over 2.6 M samples), two orders below the guard band that flags unreliable signs.  Row 0
// of every lane runs in the reference's order (the offset term of the first six samples of a stream differs) and
// its last six outputs give the all-pole states: w3 = y, w2 = A3 w3, w1 = A2 w2.
// BULK: the rows are staged by the TMA unit instead of the LSU -- every lane issues ONE 128-byte cp.async.bulk
// (UBLKCP) per row, straight into its own staging row, completing on the stage's mbarrier (lane 0 announces the
// bytes with arrive.expect_tx, all lanes wait on the phase parity) -- in place of eight 16-byte LDGSTS per lane with
// their shuffled source addresses and commit / wait groups.
// F64: the drop's samples are the halved double-precision signal (w.xf, recordings above 50 kHz) instead of int16:
// every lane reads its own 512-byte row straight from global memory, two samples per 16-byte load (whole lines are
// consumed, and the next row's four lines are prefetched), because staging 512 bytes per lane and stage would not
// leave room for two CTAs per SM; everything after the sample fetch is the same code.
template <int NSEC, int NPCM, bool HEAD, bool FAST, bool BULK = false, bool F64 = false>
__device__ __forceinline__ void ax_demod_fused_body(const AxWave& w, const AxWinTab& tab, int cfg_id, int64_t n_items) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AxFdSmem& smem = *reinterpret_cast<AxFdSmem*>(ax_smem_raw);
    AxFdWarp& sm = smem.wp[warp];
    const int64_t seg = (int64_t)blockIdx.x * AX_FD_THREADS + threadIdx.x;
    int d;
    bool active, kept = false;
    AxSegGeom g;
    g.seg_start = g.seg_end = g.n_begin = g.n_stop = 0;
    int skip = 0;                                       // HEAD: samples of row 0 that lie before the chunk start
    int64_t chunk_s = 0;
    int head_k = -1;                                    // HEAD: run() iteration of this lane
    if (!HEAD) {
        d = w.seg_drop[(int64_t)blockIdx.x * AX_FD_THREADS];
        const AxDrop& dr0 = w.drop[d];
        if (dr0.cfg != cfg_id || (dr0.xf_off >= 0) != F64) return;   // another launch handles this rate class / the halved signals
        const int64_t j = seg - dr0.seg_base;
        // streaming: segments whose records an earlier run of the growing recording left final are skipped
        kept = w.streaming && j < w.st[d].seg_done;
        if (w.streaming && (int64_t)blockIdx.x * AX_FD_THREADS + AX_FD_THREADS - 1 - dr0.seg_base < w.st[d].seg_done) return;
        active = j < dr0.nseg && !kept;
        if (active) g = ax_seg_geom(dr0, w.cfg[cfg_id], w.seg_len, j);
    } else {
        const int64_t cg = seg < n_items ? seg : n_items - 1;
        d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
        const AxDrop& dr0 = w.drop[d];
        const int k = (int)(cg - dr0.chunk_base);
        head_k = k;
        active = seg < n_items && k < dr0.chunk_cap && dr0.cfg == cfg_id && dr0.xf_off < 0;
        if (active) {
            const AxHeadGeom hg = ax_head_geom(w, dr0, w.st[d], w.cfg[cfg_id], w.chunk[cg], k);
            active = hg.active;
            if (active) {
                chunk_s = hg.s;
                g.n_begin = hg.s & ~(int64_t)7; skip = (int)(hg.s - g.n_begin);
                g.n_stop = hg.s + hg.ny;
                g.seg_start = hg.s + w.cfg[cfg_id].pad; g.seg_end = hg.s + hg.H - 1;
            }
        }
        if (!__any_sync(0xffffffffu, active)) { if (seg < n_items && !active) { /* nothing to do for this warp */ } return; }
    }
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[cfg_id];
    AxState& st = w.st[d];
    const int T = active ? (int)((g.n_stop - g.n_begin + 63) >> 6) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    const int16_t* xdrop = w.pcm + dr.pcm_off;
    const unsigned long long xrow = (unsigned long long)(xdrop + g.n_begin);      // 16-byte aligned
    const double* xfrow = F64 ? w.xf + dr.xf_off + g.n_begin : nullptr;           // (n_begin and xf_off are multiples of 64)
    for (int k = lane; k < AX_WIN_TAPS; k += 32) sm.tab[k] = tab.t[k];
    // ---- cascade constants (Butterworth form, see AxFilt::filter)
    double z0[NSEC], z1[NSEC], a1[NSEC], a2[NSEC], sg[NSEC];
#pragma unroll
    for (int s = 0; s < NSEC; ++s) {
        z0[s] = 0.0; z1[s] = 0.0;
        a1[s] = -c.sos[s][4]; a2[s] = -c.sos[s][5];
        sg[s] = (c.sos[s][1] < 0.0) ? -2.0 : 2.0;
    }
    const double k0 = c.sos[0][0] * st.inv_ampl, k1 = c.sos[0][0] * -(st.dc * st.inv_ampl);
    const double k1_64 = 64.0 * k1;
    int3 hist = make_int3(0, 0, 0);         // FAST: the last three packed words of the previous row
    const float guard_f = (float)w.guard;
    const int nb = (int)g.n_begin, nstop = (int)g.n_stop, sstart = (int)g.seg_start, send = (int)g.seg_end;
    const int out_cap = HEAD ? w.head_zc_cap_max : w.seg_cap;
    const int64_t wslot0 = (seg - lane) * (int64_t)out_cap;          // output slot of the warp's row 0
    // rows this lane helps to stage: r = i*4 + prow, i = 0..7
    const int prow = lane >> 3, piece = lane & 7;
    unsigned long long src[BULK ? 1 : 8];
    int Tr[BULK ? 1 : 8];
    if constexpr (!BULK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            src[i] = __shfl_sync(0xffffffffu, xrow, i * 4 + prow) + (unsigned long long)piece * 16;
            Tr[i] = __shfl_sync(0xffffffffu, T, i * 4 + prow);
        }
    } else {
        if (lane == 0) { ax_mbar_init(&sm.mbar[0], 1u); ax_mbar_init(&sm.mbar[1], 1u); ax_mbar_init_fence(); }
    }
    auto issue = [&](int t, int s) {
        if constexpr (F64) {
            if (t < T) {
#pragma unroll
                for (int l = 0; l < 4; ++l) asm volatile("prefetch.global.L1 [%0];" ::"l"(xfrow + 64 * (long long)t + 16 * l));
            }
        } else if constexpr (BULK) {
            const bool mine = t < T;
            const unsigned nact = (unsigned)__popc(__ballot_sync(0xffffffffu, mine));
            if (lane == 0) ax_mbar_expect_tx(&sm.mbar[s], 128u * nact);
            if (mine) ax_bulk_g2s(&sm.stage[s][lane * AX_FD_ROW], reinterpret_cast<const void*>(xrow + (unsigned long long)t * 128), 128u, &sm.mbar[s]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (t < Tr[i]) ax_cp_async16(&sm.stage[s][(i * 4 + prow) * AX_FD_ROW + piece * 8],
                                             reinterpret_cast<const void*>(src[i] + (unsigned long long)t * 128));
            ax_cp_async_commit();
        }
    };
    unsigned long long Sprev = 0ull;        // sign bits of row t-1 (bit i = sample i negative)
    int count = 0, unc = 0;
    float4* myring = smem.yring[warp] + lane;        // quad q of this lane's ring: myring[32 * q]
    const unsigned ring_hi = (unsigned)warp << 14, ring_lo = (unsigned)lane << 4;
    __syncwarp();
    if (Tmax > 0) issue(0, 0);
    for (int t = 0; t <= Tmax; ++t) {
        unsigned long long Scur = 0ull;
        if (t < Tmax) {
            if constexpr (F64) {
                if (t + 1 < Tmax) issue(t + 1, 0);
            } else if constexpr (BULK) {
                if (t + 1 < Tmax) issue(t + 1, (t + 1) & 1);
                ax_mbar_wait(&sm.mbar[t & 1], (unsigned)((t >> 1) & 1));
            } else {
                if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
            }
            __syncwarp();
            // ---------------- phase 1: the cascade over this lane's 64 samples
            if (t < T) {
                const int4* rp = reinterpret_cast<const int4*>(&sm.stage[t & 1][lane * AX_FD_ROW]);
                float4* yo = myring + (t & 1) * (16 * 32);
                float minabs = 1e30f;
                unsigned sb = 0u;
                float yf[4];
                auto emit = [&](const int m, const double y) {
                    sb = __funnelshift_l((unsigned)__double2hiint(y), sb, 1);   // MSB-first: sample 0 of this half ends at bit 31
                    const float f = (float)y;
                    yf[m & 3] = f;
                    minabs = fminf(minabs, fabsf(f));
                    if ((m & 3) == 3) yo[32 * (m >> 2)] = make_float4(yf[0], yf[1], yf[2], yf[3]);
                    if ((m & 31) == 31) { Scur |= (unsigned long long)__brev(sb) << (m & 32); sb = 0u; }
                };
                // The sections run skewed by one sample each (section s works on sample n - s), so every step
                // holds NSEC independent recurrences.
                double pipe[NSEC];                   // pipe[s]: output of section s-1 for the sample section s takes next
                int4 q = F64 ? make_int4(0, 0, 0, 0) : rp[0];
                double2 xq = make_double2(0.0, 0.0);
                if (!FAST || t == 0) {
                    // per sample the arithmetic is exactly AxFilt::filter's
                    double yl[6];                    // FAST: outputs 58..63 of the row
#pragma unroll
                    for (int n = 0; n < 64 + NSEC - 1; ++n) {
#pragma unroll
                        for (int s = NSEC - 1; s >= 0; --s) {
                            const int m = n - s;         // sample this section handles in this step
                            if (m >= 0 && m < 64) {
                                double tt;
                                if (s == 0) {
                                    if constexpr (F64) {
                                        if ((m & 1) == 0) xq = __ldg(reinterpret_cast<const double2*>(xfrow + 64 * (long long)t + m));
                                        tt = fma((m & 1) ? xq.y : xq.x, k0, k1);
                                    } else {
                                    if ((m & 7) == 0 && m > 0) q = rp[m >> 3];
                                    const int wdv = ((m & 7) >> 1) == 0 ? q.x : ((m & 7) >> 1) == 1 ? q.y : ((m & 7) >> 1) == 2 ? q.z : q.w;
                                    const short2 xs2 = *reinterpret_cast<const short2*>(&wdv);   // I2F.F64.S16 on either half
                                    tt = fma((m & 1) ? (double)xs2.y : (double)xs2.x, k0, k1);
                                    if (HEAD && t == 0 && m < skip) tt = 0.0;      // before the chunk start: keeps the state at zero
                                    }
                                } else tt = pipe[s];
                                const double y = tt + z0[s];
                                z0[s] = fma(a1[s], y, fma(sg[s], tt, z1[s]));
                                z1[s] = fma(a2[s], y, tt);
                                if (s < NSEC - 1) pipe[s + 1] = y;
                                else {
                                    if (FAST && m >= 58) yl[m - 58] = y;
                                    emit(m, y);
                                }
                            }
                        }
                    }
                    if (FAST) {
                        // all-pole states after sample 63: z0[s] = y_s[63], z1[s] = y_s[62] (y_s = A_{s+1} y_{s+1}, y_2 = y)
                        double w2[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) w2[i] = fma(-a1[2], yl[i + 1], fma(-a2[2], yl[i], yl[i + 2]));     // samples 60..63
                        z0[2] = yl[5]; z1[2] = yl[4];
                        z0[1] = w2[3]; z1[1] = w2[2];
                        z0[0] = fma(-a1[1], w2[2], fma(-a2[1], w2[1], w2[3]));
                        z1[0] = fma(-a1[1], w2[1], fma(-a2[1], w2[0], w2[2]));
                    }
                } else {
                    int4 qp = make_int4(0, hist.x, hist.y, hist.z);          // words -4..-1 of the row (word -4 is not used)
#pragma unroll
                    for (int n = 0; n < 64 + NSEC - 1; ++n) {
#pragma unroll
                        for (int s = NSEC - 1; s >= 0; --s) {
                            const int m = n - s;
                            if (m >= 0 && m < 64) {
                                double tt;
                                if (s == 0) {
                                    if ((m & 7) == 0 && m > 0) { qp = q; q = rp[m >> 3]; }
                                    // packed words W[k - i], k = m >> 1 (x[2k] low half, x[2k+1] high half)
                                    const int c = (m >> 1) & 3;
                                    const int cw[4] = {q.x, q.y, q.z, q.w}, pw[4] = {qp.x, qp.y, qp.z, qp.w};
                                    const int w0 = cw[c];
                                    const int w1 = c >= 1 ? cw[c - 1] : pw[c + 3];
                                    const int w2 = c >= 2 ? cw[c - 2] : pw[c + 2];
                                    const int w3 = c >= 3 ? cw[c - 3] : pw[c + 1];
                                    int u;
                                    if ((m & 1) == 0) u = __dp2a_lo(w3, 0x0601, __dp2a_lo(w2, 0x140F, __dp2a_lo(w1, 0x060F, __dp2a_lo(w0, 0x0001, 0))));
                                    else u = __dp2a_lo(w3, 0x0100, __dp2a_lo(w2, 0x0F06, __dp2a_lo(w1, 0x0F14, __dp2a_lo(w0, 0x0106, 0))));
                                    tt = fma((double)u, k0, k1_64);
                                } else tt = pipe[s];
                                const double y = fma(a1[s], z0[s], fma(a2[s], z1[s], tt));
                                z1[s] = z0[s]; z0[s] = y;
                                if (s < NSEC - 1) pipe[s + 1] = y;
                                else emit(m, y);
                            }
                        }
                    }
                }
                if (FAST) { const int4 ql = rp[7]; hist = make_int3(ql.y, ql.z, ql.w); }
                // guard band: a filter output this close to zero cannot be signed reliably (AXCTD_DROP_UNCERTAIN)
                if (minabs < guard_f) {
                    const int base = nb + 64 * t;
                    for (int i = 0; i < 64; ++i) {
                        const int n = base + i;
                        const float yv = reinterpret_cast<const float*>(&yo[32 * (i >> 2)])[i & 3];
                        if (n >= sstart && n < send && n < nstop && fabsf(yv) < guard_f) {
                            ++unc;
                            ax_unc_push(w, d, n, (__float_as_uint(yv) >> 31) != 0u, HEAD ? head_k + 1 : 0, chunk_s);
                        }
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
#ifdef AX_DEMOD_PROBE
                if (w.probe == 2) X = 0ull;
#endif
            }
            // every lane works through the crossings of its own row (its own ring: no exchange, no barrier);
            // the warp iterates as often as its busiest lane has crossings
            unsigned Xl = (unsigned)X, Xh = (unsigned)(X >> 32);
            while (__any_sync(0xffffffffu, (Xl | Xh) != 0u)) {
                const bool has = (Xl | Xh) != 0u;
                const bool in_lo = Xl != 0u;
                const unsigned xs = in_lo ? Xl : Xh;                 // (branch-free 64-bit find-first-set)
                const int p = has ? __ffs((int)xs) - 1 + (in_lo ? 0 : 32) : 0;
                const unsigned xc = xs & (xs - 1u);
                if (in_lo) Xl = xc; else Xh = xc;
                const int j0 = ((t - 1) & 1) * 64 + p + 1;           // ring position of the first window sample
                const int o = j0 & 3;
                constexpr int NQ = (NPCM + 6) >> 2;
                float yv[NQ * 4];
                const unsigned q0b = ((unsigned)(j0 >> 2) << 9) + ring_lo;
#pragma unroll
                for (int k = 0; k < NQ; ++k) {
                    const float4 v = *reinterpret_cast<const float4*>(ax_smem_raw + ((((q0b + 512u * k) & 0x3ff0u) | ring_hi)));
                    yv[4 * k] = v.x; yv[4 * k + 1] = v.y; yv[4 * k + 2] = v.z; yv[4 * k + 3] = v.w;
                }
#ifdef AX_DEMOD_PROBE
                float m1 = yv[0], m2 = yv[NQ * 4 - 1];
                if (w.probe != 1) ax_window32(yv, o, NPCM, sm.tab, &m1, &m2);
#else
                float m1, m2;
                ax_window32(yv, o, NPCM, sm.tab, &m1, &m2);
#endif
                if (has) {
                    const int idx = base + p;
                    const bool complete = idx + NPCM < nstop;
                    if (count < out_cap) {
                        const int64_t oi = wslot0 + (int64_t)lane * out_cap + count;
                        if (!HEAD) {
                            w.rec_idx[oi] = idx;
                            w.rec_a1[oi] = complete ? m1 : __int_as_float(0x7fc00000);
                            w.rec_a2[oi] = complete ? m2 : __int_as_float(0x7fc00000);
                        } else {
                            w.head_idx[oi] = idx - (int)chunk_s;             // chunk-relative
                            w.head_a1[oi] = complete ? m1 : __int_as_float(0x7fc00000);
                            w.head_a2[oi] = complete ? m2 : __int_as_float(0x7fc00000);
                        }
                    }
                    ++count;
                }
            }
        }
        Sprev = Scur;
        __syncwarp();
    }
    if (!HEAD) {
        if (active) {
            if (count > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; ax_raise(w.st[d], AXCTD_DROP_CAPACITY, -1); count = w.seg_cap; }   // crossings were dropped: fail the drop
            w.seg_cnt[seg] = count;
            w.seg_unc[seg] = unc;
        } else if (!kept) { w.seg_cnt[seg] = 0; w.seg_unc[seg] = 0; }
    } else if (active) {
        w.head_cnt[seg] = count > out_cap ? -1 : count;
        w.head_unc[seg] = unc;
    }
}

template <int NSEC, int NPCM, bool HEAD, bool FAST, bool BULK = false, bool F64 = false>
__global__ void __launch_bounds__(AX_FD_THREADS, 2) k_demod_fused(const __grid_constant__ AxWave w, const __grid_constant__ AxWinTab tab, int cfg_id, int64_t n_items) {
    ax_demod_fused_body<NSEC, NPCM, HEAD, FAST, BULK, F64>(w, tab, cfg_id, n_items);
}
// Both rate classes of a batch (window lengths 39 and 43 samples: 44.1 and 48 kHz) in ONE launch: a CTA takes the body
// of its drop's class.  The segment length of a batch is chosen so that ALL its segments fill a whole number of waves
// of (SMs x 8 warps x 32) lanes; one launch per class broke that fit (2.1 + 1.9 waves ran as 3 + 2 rounds).
template <int NSEC, bool FAST>
__global__ void __launch_bounds__(AX_FD_THREADS, 2) k_demod_fused_pair(const __grid_constant__ AxWave w, const __grid_constant__ AxWinTab tab39, const __grid_constant__ AxWinTab tab43,
                                                                        int cfg39, int cfg43, int64_t n_items) {
    const int cfg = w.drop[w.seg_drop[(int64_t)blockIdx.x * AX_FD_THREADS]].cfg;
    if (cfg == cfg39) ax_demod_fused_body<NSEC, 39, false, FAST>(w, tab39, cfg39, n_items);
    else if (cfg == cfg43) ax_demod_fused_body<NSEC, 43, false, FAST>(w, tab43, cfg43, n_items);
}

// ------------------------------------------------------------------ fused demodulation pass, warp-specialised
// k_demod_ws: the computation of k_demod_fused split over two kinds of warps so that the FP64 cascade and the
// FP32 windows overlap instead of alternating.  A CTA holds AX_WS_PAIRS pairs of warps; pair p owns 32
// segments (lane = segment in both warps):
//   filter warp  stages its lanes' int16 rows (32 samples, 16-byte cp.async, double buffered), runs the skewed
//                SOS cascade, stores the float roundings of y into the pair's ring in shared memory
//                ([quad][lane], four rows deep) and the row's sign word, then arrives on full[row & 3];
//   window warp  waits on full[u & 3], takes the crossings of row u-2 (their windows end inside row u), sums
//                the windows from the ring exactly as k_demod_fused does, writes the crossing records and
//                arrives on free[(u-2) & 3], which the filter warp waits on before it overwrites that slot.
// The filter warp may run one row ahead of the window warp; with two CTAs per SM every scheduler holds two
// filter warps (six independent DFMA chains) and two window warps.
#define AX_WS_PAIRS 4
#define AX_WS_THREADS (AX_WS_PAIRS * 64)
#define AX_WS_RS 32                                    // samples per row
#define AX_WS_ROW 40                                   // int16 per staged row: 32 samples + 8 pad (80-byte stride: conflict-free LDS.128)
#define AX_WS_STAGE (32 * AX_WS_ROW)
#define AX_WS_SLOTS 4                                  // ring rows per lane
#define AX_WS_RINGQ (AX_WS_SLOTS * AX_WS_RS / 4)       // quads per lane ring
struct AxWsPair {
    int16_t stage[2][AX_WS_STAGE];
    float4 yring[AX_WS_RINGQ * 32];                    // [quad][lane]
    uint32_t signs[AX_WS_SLOTS][32];                   // bit i = sample i of the row is negative
    unsigned long long full[AX_WS_SLOTS], freeb[AX_WS_SLOTS];
};
struct AxWsSmem {
    AxF4 tab[AX_WIN_TAPS];
    AxWsPair pr[AX_WS_PAIRS];
};
template <int NSEC, int NPCM, bool HEAD>
__global__ void __launch_bounds__(AX_WS_THREADS, 2) k_demod_ws(const __grid_constant__ AxWave w, const __grid_constant__ AxWinTab tab, int cfg_id, int64_t n_items) {
    static_assert(NPCM + 2 <= 2 * AX_WS_RS, "window must end inside the row after next");
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp % AX_WS_PAIRS;
    const bool is_filter = warp < AX_WS_PAIRS;
    AxWsSmem& smem = *reinterpret_cast<AxWsSmem*>(ax_smem_raw);
    AxWsPair& sm = smem.pr[pair];
    const int64_t seg0 = (int64_t)blockIdx.x * (AX_WS_PAIRS * 32);
    const int64_t seg = seg0 + pair * 32 + lane;
    int d;
    bool active;
    AxSegGeom g;
    g.seg_start = g.seg_end = g.n_begin = g.n_stop = 0;
    int skip = 0;                                       // HEAD: samples of row 0 that lie before the chunk start
    int64_t chunk_s = 0;
    int head_k = -1;
    if (!HEAD) {
        d = w.seg_drop[seg0];
        const AxDrop& dr0 = w.drop[d];
        if (dr0.cfg != cfg_id || dr0.xf_off >= 0) return;   // another launch handles this rate class / the generic kernel the halved signals
        const int64_t j = seg - dr0.seg_base;
        active = j < dr0.nseg;
        if (active) g = ax_seg_geom(dr0, w.cfg[cfg_id], w.seg_len, j);
    } else {
        const int64_t cg = seg < n_items ? seg : n_items - 1;
        d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
        const AxDrop& dr0 = w.drop[d];
        const int k = (int)(cg - dr0.chunk_base);
        head_k = k;
        active = seg < n_items && k < dr0.chunk_cap && dr0.cfg == cfg_id && dr0.xf_off < 0;
        if (active) {
            const AxHeadGeom hg = ax_head_geom(w, dr0, w.st[d], w.cfg[cfg_id], w.chunk[cg], k);
            active = hg.active;
            if (active) {
                chunk_s = hg.s;
                g.n_begin = hg.s & ~(int64_t)7; skip = (int)(hg.s - g.n_begin);
                g.n_stop = hg.s + hg.ny;
                g.seg_start = hg.s + w.cfg[cfg_id].pad; g.seg_end = hg.s + hg.H - 1;
            }
        }
        if (__syncthreads_or(active ? 1 : 0) == 0) return;          // no head in this CTA
    }
    const AxCfg& c = w.cfg[cfg_id];
    AxState& st = w.st[d];
    for (int k = threadIdx.x; k < AX_WIN_TAPS; k += AX_WS_THREADS) smem.tab[k] = tab.t[k];
    if (is_filter && lane < 2 * AX_WS_SLOTS) ax_mbar_init(lane < AX_WS_SLOTS ? &sm.full[lane] : &sm.freeb[lane - AX_WS_SLOTS], 32u);
    const int T = active ? (int)((g.n_stop - g.n_begin + AX_WS_RS - 1) / AX_WS_RS) : 0;
    int Tmax = T;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
    const int nb = (int)g.n_begin, nstop = (int)g.n_stop, sstart = (int)g.seg_start, send = (int)g.seg_end;
    float4* myring = sm.yring + lane;                // quad q of this lane's ring: myring[32 * q]
    __syncthreads();
    if (is_filter) {
        // ------------------------------------------------------------ filter warp
        const unsigned long long xrow = (unsigned long long)(w.pcm + w.drop[d].pcm_off + g.n_begin);   // 16-byte aligned
        double z0[NSEC], z1[NSEC], a1[NSEC], a2[NSEC], sg[NSEC];
#pragma unroll
        for (int s = 0; s < NSEC; ++s) {
            z0[s] = 0.0; z1[s] = 0.0;
            a1[s] = -c.sos[s][4]; a2[s] = -c.sos[s][5];
            sg[s] = (c.sos[s][1] < 0.0) ? -2.0 : 2.0;
        }
        const double k0 = c.sos[0][0] * st.inv_ampl, k1 = c.sos[0][0] * -(st.dc * st.inv_ampl);
        const float guard_f = (float)w.guard;
        // rows this lane helps to stage: r = i*8 + prow, i = 0..3 (four 16-byte pieces per row)
        const int prow = lane >> 2, piece = lane & 3;
        unsigned long long src[4];
        int Tr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            src[i] = __shfl_sync(0xffffffffu, xrow, i * 8 + prow) + (unsigned long long)piece * 16;
            Tr[i] = __shfl_sync(0xffffffffu, T, i * 8 + prow);
        }
        auto issue = [&](int t, int s) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t < Tr[i]) ax_cp_async16(&sm.stage[s][(i * 8 + prow) * AX_WS_ROW + piece * 8],
                                             reinterpret_cast<const void*>(src[i] + (unsigned long long)t * (2 * AX_WS_RS)));
            ax_cp_async_commit();
        };
        int unc = 0;
        if (Tmax > 0) issue(0, 0);
        for (int t = 0; t < Tmax; ++t) {
            if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
            __syncwarp();
            if (t >= AX_WS_SLOTS) ax_mbar_wait(&sm.freeb[t & (AX_WS_SLOTS - 1)], (unsigned)(((t - AX_WS_SLOTS) / AX_WS_SLOTS) & 1));
            unsigned S = 0u;
            if (t < T) {
                const int4* rp = reinterpret_cast<const int4*>(&sm.stage[t & 1][lane * AX_WS_ROW]);
                float4* yo = myring + (t & (AX_WS_SLOTS - 1)) * (8 * 32);
                float minabs = 1e30f;
                // The sections run skewed by one sample each (section s works on sample n - s), so every step
                // holds NSEC independent recurrences; per sample the arithmetic is exactly AxFilt::filter's.
                double pipe[NSEC];
                int4 q = rp[0];
                unsigned sb = 0u;
                float yf[4];
#pragma unroll
                for (int n = 0; n < AX_WS_RS + NSEC - 1; ++n) {
#pragma unroll
                    for (int s = NSEC - 1; s >= 0; --s) {
                        const int m = n - s;         // sample this section handles in this step
                        if (m >= 0 && m < AX_WS_RS) {
                            double tt;
                            if (s == 0) {
                                if ((m & 7) == 0 && m > 0) q = rp[m >> 3];
                                const int wdv = ((m & 7) >> 1) == 0 ? q.x : ((m & 7) >> 1) == 1 ? q.y : ((m & 7) >> 1) == 2 ? q.z : q.w;
                                const int xi = (m & 1) ? (wdv >> 16) : (int)(short)(wdv & 0xFFFF);
                                tt = fma((double)xi, k0, k1);
                                if (HEAD && t == 0 && m < skip) tt = 0.0;      // before the chunk start: keeps the state at zero
                            } else tt = pipe[s];
                            const double y = tt + z0[s];
                            z0[s] = fma(a1[s], y, fma(sg[s], tt, z1[s]));
                            z1[s] = fma(a2[s], y, tt);
                            if (s < NSEC - 1) pipe[s + 1] = y;
                            else {
                                sb = __funnelshift_l((unsigned)__double2hiint(y), sb, 1);   // MSB-first: sample 0 ends at bit 31
                                const float f = (float)y;
                                yf[m & 3] = f;
                                minabs = fminf(minabs, fabsf(f));
                                if ((m & 3) == 3) yo[32 * (m >> 2)] = make_float4(yf[0], yf[1], yf[2], yf[3]);
                            }
                        }
                    }
                }
                S = __brev(sb);
                // guard band: a filter output this close to zero cannot be signed reliably (AXCTD_DROP_UNCERTAIN)
                if (minabs < guard_f) {
                    const int base = nb + AX_WS_RS * t;
                    for (int i = 0; i < AX_WS_RS; ++i) {
                        const int n = base + i;
                        const float yv = reinterpret_cast<const float*>(&yo[32 * (i >> 2)])[i & 3];
                        if (n >= sstart && n < send && n < nstop && fabsf(yv) < guard_f) {
                            ++unc;
                            ax_unc_push(w, d, n, (__float_as_uint(yv) >> 31) != 0u, HEAD ? head_k + 1 : 0, chunk_s);
                        }
                    }
                }
            }
            sm.signs[t & (AX_WS_SLOTS - 1)][lane] = S;
            ax_mbar_arrive(&sm.full[t & (AX_WS_SLOTS - 1)]);
            __syncwarp();                                // the stage buffer t & 1 is rewritten by issue(t + 2)
        }
        if (!HEAD) w.seg_unc[seg] = active ? unc : 0;
        else if (active) w.head_unc[seg] = unc;
        return;
    }
    // ---------------------------------------------------------------- window warp
    const int out_cap = HEAD ? w.head_zc_cap_max : w.seg_cap;
    const int64_t oslot = seg * (int64_t)out_cap;
    int count = 0;
    for (int u = 2; u < Tmax + 2; ++u) {
        const int r = u - 2;                            // row whose crossings are due
        {
            const int rw = min(u, Tmax - 1);            // newest row the windows of row r can reach
            ax_mbar_wait(&sm.full[rw & (AX_WS_SLOTS - 1)], (unsigned)((rw / AX_WS_SLOTS) & 1));
        }
        unsigned X = 0u;
        const int base = nb + AX_WS_RS * r;
        if (r < T) {
            const unsigned S = sm.signs[r & (AX_WS_SLOTS - 1)][lane];
            const unsigned nxt = (r + 1 < T) ? (sm.signs[(r + 1) & (AX_WS_SLOTS - 1)][lane] & 1u) : (S >> 31);
            X = S ^ ((S >> 1) | (nxt << 31));
            int lo = sstart - base, hi = min(send, nstop - 1) - base;      // crossing i needs sample i+1
            lo = max(lo, 0); hi = min(hi, 32);
            unsigned m = 0u;
            if (hi > lo) m = ((hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u)) & ~((1u << lo) - 1u);
            X &= m;
        }
        // every lane works through the crossings of its own row (its own ring column: no exchange);
        // the warp iterates as often as its busiest lane has crossings
        while (__any_sync(0xffffffffu, X != 0u)) {
            const bool has = X != 0u;
            const int p = has ? __ffs((int)X) - 1 : 0;
            if (has) X &= X - 1u;
            const int j0 = (r & (AX_WS_SLOTS - 1)) * AX_WS_RS + p + 1;     // ring position of the first window sample
            const int o = j0 & 3;
            constexpr int NQ = (NPCM + 6) >> 2;
            float yv[NQ * 4];
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
                const float4 v = myring[32 * (((j0 >> 2) + k) & (AX_WS_RINGQ - 1))];
                yv[4 * k] = v.x; yv[4 * k + 1] = v.y; yv[4 * k + 2] = v.z; yv[4 * k + 3] = v.w;
            }
            float m1, m2;
            ax_window32(yv, o, NPCM, smem.tab, &m1, &m2);
            if (has) {
                const int idx = base + p;
                const bool complete = idx + NPCM < nstop;
                if (count < out_cap) {
                    const int64_t oi = oslot + count;
                    if (!HEAD) {
                        w.rec_idx[oi] = idx;
                        w.rec_a1[oi] = complete ? m1 : __int_as_float(0x7fc00000);
                        w.rec_a2[oi] = complete ? m2 : __int_as_float(0x7fc00000);
                    } else {
                        w.head_idx[oi] = idx - (int)chunk_s;             // chunk-relative
                        w.head_a1[oi] = complete ? m1 : __int_as_float(0x7fc00000);
                        w.head_a2[oi] = complete ? m2 : __int_as_float(0x7fc00000);
                    }
                }
                ++count;
            }
        }
        ax_mbar_arrive(&sm.freeb[r & (AX_WS_SLOTS - 1)]);
    }
    if (!HEAD) {
        if (active) {
            if (count > w.seg_cap) { w.flags[AX_FLAG_CAP] = 1; ax_raise(w.st[d], AXCTD_DROP_CAPACITY, -1); count = w.seg_cap; }   // crossings were dropped: fail the drop
            w.seg_cnt[seg] = count;
        } else w.seg_cnt[seg] = 0;
    } else if (active) {
        w.head_cnt[seg] = count > out_cap ? -1 : count;
    }
}

template <int NSEC, int NPCM, bool HEAD>
static inline void ax_launch_demod_ws(const AxWave& w, const AxCfg& c, int cfg_id, int64_t n_items, cudaStream_t stream, int device) {
    const size_t smem = sizeof(AxWsSmem);
    ax_optin_smem<k_demod_ws<NSEC, NPCM, HEAD>>(smem, device);
    const int64_t items = HEAD ? n_items : (int64_t)w.nseg_total;
    if (items <= 0) return;
    const int per = AX_WS_PAIRS * 32;
    k_demod_ws<NSEC, NPCM, HEAD><<<(unsigned)((items + per - 1) / per), AX_WS_THREADS, smem, stream>>>(w, c.win_tab, cfg_id, items);
}

template <int NSEC, int NPCM, bool HEAD, bool FAST = false, bool BULK = false, bool F64 = false>
static inline void ax_launch_demod_fused(const AxWave& w, const AxCfg& c, int cfg_id, int64_t n_items, cudaStream_t stream, int device) {
    const size_t smem = sizeof(AxFdSmem);
    ax_optin_smem<k_demod_fused<NSEC, NPCM, HEAD, FAST, BULK, F64>>(smem, device);
    const int64_t items = HEAD ? n_items : (int64_t)w.nseg_total;
    if (items <= 0) return;
    k_demod_fused<NSEC, NPCM, HEAD, FAST, BULK, F64><<<(unsigned)((items + AX_FD_THREADS - 1) / AX_FD_THREADS), AX_FD_THREADS, smem, stream>>>(w, c.win_tab, cfg_id, items);
}
// continuous pass over the halved (double-precision) signals of rate class cfg_id
static inline void ax_launch_demod_fused_f64(const AxWave& w, const AxCfg& c, int cfg_id, cudaStream_t stream, int device) {
    if (c.nsec == 3 && c.npcm == 39) ax_launch_demod_fused<3, 39, false, false, false, true>(w, c, cfg_id, 0, stream, device);
    else if (c.nsec == 3 && c.npcm == 43) ax_launch_demod_fused<3, 43, false, false, false, true>(w, c, cfg_id, 0, stream, device);
    else if (c.nsec == 6 && c.npcm == 39) ax_launch_demod_fused<6, 39, false, false, false, true>(w, c, cfg_id, 0, stream, device);
    else ax_launch_demod_fused<6, 43, false, false, false, true>(w, c, cfg_id, 0, stream, device);
}

// continuous pass of two low-pass rate classes (npcm 39 and 43) in one launch
template <int NSEC, bool FAST>
static inline void ax_launch_demod_fused_pair(const AxWave& w, const AxCfg& c39, int cfg39, const AxCfg& c43, int cfg43, cudaStream_t stream, int device) {
    const size_t smem = sizeof(AxFdSmem);
    ax_optin_smem<k_demod_fused_pair<NSEC, FAST>>(smem, device);
    const int64_t items = (int64_t)w.nseg_total;
    if (items <= 0) return;
    k_demod_fused_pair<NSEC, FAST><<<(unsigned)((items + AX_FD_THREADS - 1) / AX_FD_THREADS), AX_FD_THREADS, smem, stream>>>(w, c39.win_tab, c43.win_tab, cfg39, cfg43, items);
}
static inline bool ax_demod_fast_ok(const AxCfg& c) {       // numerators-first cascade: low-pass, every section with numerator g (1 + z^-1)^2
    return c.nsec == 3 && c.sos[0][1] > 0.0 && c.sos[1][1] > 0.0 && c.sos[2][1] > 0.0;
}

// true if the fused kernel has an instantiation for this rate class
static inline bool ax_demod_fused_ok(const AxCfg& c) {
    return ax_sos_is_butter(c) && (c.nsec == 3 || c.nsec == 6) && (c.npcm == 39 || c.npcm == 43) && c.inset == 1;
}
template <bool HEAD>
static inline void ax_launch_demod_fused_any(const AxWave& w, const AxCfg& c, int cfg_id, int64_t n_items, cudaStream_t stream, int device, int ws, int fast, int bulk) {
    // numerators-first cascade: continuous pass of the low-pass (every section with numerator g (1 + z^-1)^2) only
    if (!HEAD && !ws && fast && c.nsec == 3 && c.sos[0][1] > 0.0 && c.sos[1][1] > 0.0 && c.sos[2][1] > 0.0) {
        if (bulk) {
            if (c.npcm == 39) ax_launch_demod_fused<3, 39, false, true, true>(w, c, cfg_id, n_items, stream, device);
            else ax_launch_demod_fused<3, 43, false, true, true>(w, c, cfg_id, n_items, stream, device);
        } else {
            if (c.npcm == 39) ax_launch_demod_fused<3, 39, false, true>(w, c, cfg_id, n_items, stream, device);
            else ax_launch_demod_fused<3, 43, false, true>(w, c, cfg_id, n_items, stream, device);
        }
        return;
    }
    if (!HEAD && !ws && bulk) {                          // reference-order cascade (band-pass, fir_first = 0) with bulk staging
        if (c.nsec == 3 && c.npcm == 39) ax_launch_demod_fused<3, 39, false, false, true>(w, c, cfg_id, n_items, stream, device);
        else if (c.nsec == 3 && c.npcm == 43) ax_launch_demod_fused<3, 43, false, false, true>(w, c, cfg_id, n_items, stream, device);
        else if (c.nsec == 6 && c.npcm == 39) ax_launch_demod_fused<6, 39, false, false, true>(w, c, cfg_id, n_items, stream, device);
        else ax_launch_demod_fused<6, 43, false, false, true>(w, c, cfg_id, n_items, stream, device);
        return;
    }
    if (ws) {
        if (c.nsec == 3 && c.npcm == 39) ax_launch_demod_ws<3, 39, HEAD>(w, c, cfg_id, n_items, stream, device);
        else if (c.nsec == 3 && c.npcm == 43) ax_launch_demod_ws<3, 43, HEAD>(w, c, cfg_id, n_items, stream, device);
        else if (c.nsec == 6 && c.npcm == 39) ax_launch_demod_ws<6, 39, HEAD>(w, c, cfg_id, n_items, stream, device);
        else ax_launch_demod_ws<6, 43, HEAD>(w, c, cfg_id, n_items, stream, device);
        return;
    }
    if (c.nsec == 3 && c.npcm == 39) ax_launch_demod_fused<3, 39, HEAD>(w, c, cfg_id, n_items, stream, device);
    else if (c.nsec == 3 && c.npcm == 43) ax_launch_demod_fused<3, 43, HEAD>(w, c, cfg_id, n_items, stream, device);
    else if (c.nsec == 6 && c.npcm == 39) ax_launch_demod_fused<6, 39, HEAD>(w, c, cfg_id, n_items, stream, device);
    else ax_launch_demod_fused<6, 43, HEAD>(w, c, cfg_id, n_items, stream, device);
}

// ------------------------------------------------------------------ /2 decimation, staged (AXCTDprocessor.py:60-62)
// k_decim_fused<PASS, PAR>: the two passes of ax_decim_item (forward over the odd-extended recording -> fwd, backward
// over fwd -> every second sample of the unpadded range -> xf) with the layout of the demodulation pass: one lane per
// decimation segment, 32 samples ("a row") per iteration, rows staged by the warp with 16-byte cp.async (coalesced:
// the thread-per-segment form read 2 or 8 bytes per thread from 32 different lines per instruction and kept its
// filter state in local memory), the four Chebyshev sections skewed by one sample each so that a lane holds four
// independent recurrences, and the row's outputs transposed through shared memory so that the warp stores whole
// 256-byte (pass 0) / 128-byte (pass 1) runs.  Per sample and section the arithmetic is ax_decim_step's.  A drop's
// segments are padded to a multiple of 32 (axctd_batch_create), so a warp never spans two drops and the geometry of
// lane q is lane 0's shifted by q segments; rows that touch the padding of the extension or the ends of the
// recording (first / last rows of a drop) take a per-sample path that reads global memory directly.
#define AX_DC_WARPS 4
#define AX_DC_THREADS (AX_DC_WARPS * 32)
#define AX_DC_R 32
struct AxDc0Warp { int16_t stage[2][32 * 40]; double tile[32 * 33]; };      // 80-byte staging rows (conflict-free LDS.128), 33-double tile rows
struct AxDc1Warp { double stage[2][32 * 34]; double tile[32 * 17]; };       // 272-byte staging rows, 17-double tile rows
__device__ __forceinline__ long long ax_floor_div(long long a, long long b) { long long q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
__device__ __forceinline__ double ax_dc_section(double u, const double* k, double& z0, double& z1) {
    const double y = fma(k[0], u, z0);
    z0 = fma(k[1], u, fma(k[3], y, z1));
    z1 = fma(k[2], u, -__dmul_rn(k[4], y));
    return y;
}
template <int PASS, int PAR>
__global__ void __launch_bounds__(AX_DC_THREADS) k_decim_fused(const __grid_constant__ AxWave w, int64_t n_items) {
    extern __shared__ __align__(16) unsigned char ax_smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t sg0 = (int64_t)blockIdx.x * AX_DC_THREADS + warp * 32;
    if (sg0 >= n_items) return;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::dseg_base, sg0);
    const AxDrop& dr = w.drop[d];
    const long long j0 = sg0 - dr.dseg_base;
    if (dr.xf_off < 0 || j0 >= dr.ndseg) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const AxState& st = w.st[d];
    const long long N = dr.n_raw, P = c.dpad, E = N + 2 * P, DL = w.dseg_len, DR = DL / AX_DC_R;
    const long long j = j0 + lane;
    const bool active = j < dr.ndseg;
    double k[4][5], z[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        k[q][0] = c.dsos[q][0]; k[q][1] = c.dsos[q][1]; k[q][2] = c.dsos[q][2]; k[q][3] = -c.dsos[q][4]; k[q][4] = c.dsos[q][5];
        z[q][0] = 0.0; z[q][1] = 0.0;
    }
    double* fwd = w.fwd + dr.fwd_off;
    if (PASS == 0) {
        AxDc0Warp& sm = reinterpret_cast<AxDc0Warp*>(ax_smem_raw)[warp];
        const int16_t* x = w.pcm + dr.pcm_off;
        const double kmul = st.inv_ampl, kadd = -(st.dc * st.inv_ampl);
        // lane q: rows rA0 + q DR .. of 32 samples in recording coordinates n = e - P; its outputs are e in [(j0+q) DL, ...)
        const long long rA0 = ax_floor_div(j0 * DL - c.dwarm - P, AX_DC_R);
        const long long e1 = active ? ((j + 1) * DL < E ? (j + 1) * DL : E) : 0;
        const int T = active ? (int)(ax_floor_div(e1 - P + AX_DC_R - 1, AX_DC_R) - (rA0 + lane * DR)) : 0;
        int Tmax = T;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
        const int prow = lane >> 2, piece = lane & 3;
        auto issue = [&](int t, int sbuf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = i * 8 + prow;
                const long long nr = AX_DC_R * (rA0 + q * DR + t);
                if (j0 + q < dr.ndseg && nr >= 0 && nr + AX_DC_R <= N)
                    ax_cp_async16(&sm.stage[sbuf][q * 40 + piece * 8], x + nr + piece * 8);
            }
            ax_cp_async_commit();
        };
        if (Tmax > 0) issue(0, 0);
        for (int t = 0; t < Tmax; ++t) {
            if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
            __syncwarp();
            if (t < T) {
                const long long nr = AX_DC_R * (rA0 + lane * DR + t);
                double* to = sm.tile + lane * 33;
                if (nr >= 0 && nr + AX_DC_R <= N) {
                    const int4* rp = reinterpret_cast<const int4*>(&sm.stage[t & 1][lane * 40]);
                    double pipe[4];
                    int4 qd = rp[0];
#pragma unroll
                    for (int nn = 0; nn < AX_DC_R + 3; ++nn) {
#pragma unroll
                        for (int sct = 3; sct >= 0; --sct) {
                            const int m = nn - sct;
                            if (m >= 0 && m < AX_DC_R) {
                                double u;
                                if (sct == 0) {
                                    if ((m & 7) == 0 && m > 0) qd = rp[m >> 3];
                                    const int wdv = ((m & 7) >> 1) == 0 ? qd.x : ((m & 7) >> 1) == 1 ? qd.y : ((m & 7) >> 1) == 2 ? qd.z : qd.w;
                                    const short2 xs2 = *reinterpret_cast<const short2*>(&wdv);
                                    u = fma((m & 1) ? (double)xs2.y : (double)xs2.x, kmul, kadd);
                                } else u = pipe[sct];
                                const double y = ax_dc_section(u, k[sct], z[sct][0], z[sct][1]);
                                if (sct < 3) pipe[sct + 1] = y; else to[m] = y;
                            }
                        }
                    }
                } else {
                    for (int i = 0; i < AX_DC_R; ++i) {
                        const long long e = nr + i + P;
                        if (e < 0 || e >= E) continue;
                        double u = ax_decim_ext(w, dr, c, st, e);
                        if (e == 0) for (int q = 0; q < 4; ++q) { z[q][0] = ax_mul(c.dzi[q][0], u); z[q][1] = ax_mul(c.dzi[q][1], u); }
#pragma unroll
                        for (int q = 0; q < 4; ++q) u = ax_dc_section(u, k[q], z[q][0], z[q][1]);
                        to[i] = u;
                    }
                }
            }
            __syncwarp();
            const long long erow0 = AX_DC_R * (rA0 + t) + P;
#pragma unroll 4
            for (int q = 0; q < 32; ++q) {
                const long long e = erow0 + q * DL + lane, lo = (j0 + q) * DL;
                long long hi = lo + DL; if (hi > E) hi = E;
                if (j0 + q < dr.ndseg && e >= lo && e < hi) fwd[e] = sm.tile[q * 33 + lane];
            }
            __syncwarp();
        }
    } else {
        AxDc1Warp& sm = reinterpret_cast<AxDc1Warp*>(ax_smem_raw)[warp];
        double* xf = w.xf + dr.xf_off;
        // reversed segment j covers e' = E-1-e in [j DL, (j+1) DL): lane q walks rows rT0 - q DR - t of 32 samples, downwards
        const long long rT0 = ax_floor_div(E - 1 - j0 * DL + c.dwarm, AX_DC_R);
        const long long e1r = active ? ((j + 1) * DL < E ? (j + 1) * DL : E) : 0;
        const int T = active ? (int)((rT0 - lane * DR) - ax_floor_div(E - e1r, AX_DC_R) + 1) : 0;
        int Tmax = T;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Tmax = max(Tmax, __shfl_xor_sync(0xffffffffu, Tmax, o));
        const int prow = lane >> 4, piece = lane & 15;
        auto issue = [&](int t, int sbuf) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int q = i * 2 + prow;
                const long long er = AX_DC_R * (rT0 - q * DR - t);
                if (j0 + q < dr.ndseg && er >= 0 && er + AX_DC_R <= E)
                    ax_cp_async16(&sm.stage[sbuf][q * 34 + piece * 2], fwd + er + piece * 2);
            }
            ax_cp_async_commit();
        };
        if (Tmax > 0) issue(0, 0);
        for (int t = 0; t < Tmax; ++t) {
            if (t + 1 < Tmax) { issue(t + 1, (t + 1) & 1); ax_cp_async_wait<1>(); } else ax_cp_async_wait<0>();
            __syncwarp();
            if (t < T) {
                const long long er = AX_DC_R * (rT0 - lane * DR - t);
                double* to = sm.tile + lane * 17;
                if (er >= 0 && er + AX_DC_R <= E) {
                    const double2* rp = reinterpret_cast<const double2*>(&sm.stage[t & 1][lane * 34]);
                    double pipe[4];
                    double2 qd = rp[15];
#pragma unroll
                    for (int nn = 0; nn < AX_DC_R + 3; ++nn) {
#pragma unroll
                        for (int sct = 3; sct >= 0; --sct) {
                            const int m = nn - sct;
                            if (m >= 0 && m < AX_DC_R) {
                                const int i = AX_DC_R - 1 - m;         // sample of the row (descending)
                                double u;
                                if (sct == 0) {
                                    if ((i & 1) == 1 && i < AX_DC_R - 1) qd = rp[i >> 1];
                                    u = (i & 1) ? qd.y : qd.x;
                                } else u = pipe[sct];
                                const double y = ax_dc_section(u, k[sct], z[sct][0], z[sct][1]);
                                if (sct < 3) pipe[sct + 1] = y;
                                else if (((i ^ PAR) & 1) == 0) to[(i - PAR) >> 1] = y;
                            }
                        }
                    }
                } else {
                    for (int i = AX_DC_R - 1; i >= 0; --i) {
                        const long long e = er + i;
                        if (e < 0 || e >= E) continue;
                        double u = fwd[e];
                        if (e == E - 1) for (int q = 0; q < 4; ++q) { z[q][0] = ax_mul(c.dzi[q][0], u); z[q][1] = ax_mul(c.dzi[q][1], u); }
#pragma unroll
                        for (int q = 0; q < 4; ++q) u = ax_dc_section(u, k[q], z[q][0], z[q][1]);
                        if (((i ^ PAR) & 1) == 0) to[(i - PAR) >> 1] = u;
                    }
                }
            }
            __syncwarp();
#pragma unroll 4
            for (int q2 = 0; q2 < 16; ++q2) {
                const int q = 2 * q2 + (lane >> 4), uu = lane & 15;
                const long long e = AX_DC_R * (rT0 - q * DR - t) + PAR + 2 * uu, m = e - P;
                long long hi = (j0 + q + 1) * DL; if (hi > E) hi = E;
                // outputs of reversed segment j0+q: e' = E-1-e in [(j0+q) DL, hi)
                if (j0 + q < dr.ndseg && E - 1 - e >= (j0 + q) * DL && E - 1 - e < hi && m >= 0 && m < N) xf[m >> 1] = sm.tile[q * 17 + uu];
            }
            __syncwarp();
        }
    }
}
static inline bool ax_decim_fused_ok(const AxCfg& c) { return c.decimate == 2 && c.dnsec == 4; }
template <int PASS>
static inline void ax_launch_decim_fused(const AxWave& w, int64_t n_items, int par, cudaStream_t stream, int device) {
    const size_t smem = AX_DC_WARPS * (PASS == 0 ? sizeof(AxDc0Warp) : sizeof(AxDc1Warp));
    const unsigned grid = (unsigned)((n_items + AX_DC_THREADS - 1) / AX_DC_THREADS);
    if (par) { ax_optin_smem<k_decim_fused<PASS, 1>>(smem, device); k_decim_fused<PASS, 1><<<grid, AX_DC_THREADS, smem, stream>>>(w, n_items); }
    else { ax_optin_smem<k_decim_fused<PASS, 0>>(smem, device); k_decim_fused<PASS, 0><<<grid, AX_DC_THREADS, smem, stream>>>(w, n_items); }
}

// ------------------------------------------------------------------ bit decisions with shared window sums
// As ax_bits_item, one CTA per run() iteration (no per-bit searches); a window that needs double
// precision is summed by the whole warp.
// grid (chunks of the batch) for all iterations, or (iterations after the first demodulated one, drops) when only the
// first few of every drop matter (phase 0: header 1 sits within four seconds of the first pulse)
// region 0: grid = all iterations of the batch; 1: grid (iterations after k0, drops), the first nk_full of every drop;
// 2: grid = all iterations, those from k0 + nk_full on only
__global__ void __launch_bounds__(128) k_bits_chunk(AxWave w, int phase, int region) {
    int d;
    int64_t cg;
    if (region == 1) {
        d = blockIdx.y;
        cg = (int64_t)w.drop[d].chunk_base + w.st[d].k0 + blockIdx.x;
        if (w.st[d].k0 + (int)blockIdx.x >= w.drop[d].chunk_cap) return;
    } else {
        cg = blockIdx.x;
        d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
        if (region == 2 && (int)(cg - w.drop[d].chunk_base) < w.st[d].k0 + w.nk_full) return;
    }
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (st.sm_status < 1 || st.nedges_total == 0 || k < st.k0 || k >= st.n_chunks || k >= dr.chunk_cap) return;
    if (w.streaming && k < st.k_done && (phase == 1 || ax_scale_is_final(w, st))) return;     // decided by an earlier run
    const AxChunk& ch = w.chunk[cg];
    const int nb = ch.n_edges - 1;
    if (nb <= 0) return;
    if (phase == 0) {        // only the header-1 calibration window matters (same bounds as ax_bits_need)
        const AxCfg& c = w.cfg[dr.cfg];
        const int64_t mg = (int64_t)(64.0 * c.fs / c.bitrate);
        if (ch.e < st.firstpulse400 + c.h1s - c.half - mg || ch.s > st.firstpulse400 + c.h1e + c.half + mg) return;
    }
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
            ax_gwin_partial(ax_src(w, dr), f.i, f.q0, w.cfg[dr.cfg], lane, 32, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
            if (lane == L) ax_bits_fix(w, slot, f, acc);
        }
        if (phase == 1 && mine < nb) ax_bits_decide(w, d, slot);
    }
}

// The bits that k_emit_chunk (fused) listed: a warp per bit sums the double-precision window (ax_gwin_*), replaces the
// magnitudes and decides again.  The list length is only known on the device: a fixed grid strides over it.
__global__ void __launch_bounds__(128) k_bits_recheck(AxWave w) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t cnt = w.flags[AX_FLAG_FIXCNT];
    if (cnt > w.fix_cap) cnt = w.fix_cap;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < cnt; i += nwarps) {
        const int64_t slot = w.fix_list[i];
        const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::edge_base, slot);
        const AxDrop& dr = w.drop[d];
        const AxState& st = w.st[d];
        const int64_t j = slot - dr.edge_base;
        const int k = ax_chunk_of_bit(w.chunk + dr.chunk_base, st.k0, st.n_chunks, j);
        const AxChunk& ch = w.chunk[dr.chunk_base + k];
        AxBitFix f;
        f.d = d; f.i = w.edge_idx[dr.edge_base + ch.edge_off + (j - ch.bit_off)]; f.q0 = ch.s;
        double acc[4];
        ax_gwin_partial(ax_src(w, dr), f.i, f.q0, w.cfg[dr.cfg], lane, 32, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) { ax_bits_fix(w, slot, f, acc); ax_bits_decide(w, d, slot); }
    }
}

// ------------------------------------------------------------------ bit edges (CTA per run() iteration)
// ax_emit_item in parallel: head edges one per thread; the continuous part of the chunk's walk is read
// off the canonical walk tile by tile (a warp per tile, a lane per crossing: the edge number is the tile's
// running count plus the popcount below the lane's bit, no search); only the few edges stepped
// explicitly before the walk joins the canonical one are produced by one thread.
// mode 0: every iteration, magnitudes stored (the bit decisions follow in k_bits_chunk); 1: the same for the iterations
// before k0 + nk_full only; 2: the later iterations with the bit decided on the spot (ax_emit_edge, fused); 3: the later
// iterations in the two-step form (materialises the magnitudes when a caller asks for them: axctd_batch_bits)
// grid (iterations, drops): no search for the owner of an iteration; mode 1: grid (nk_full, drops), iterations counted
// from the drop's first demodulated one (a full grid spent 0.05 ms per launch on CTAs that only exit)
__global__ void __launch_bounds__(128) k_emit_chunk(AxWave w, int mode) {
    const int d = blockIdx.y;
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    const int k = (mode == 1 ? st.k0 : 0) + (int)blockIdx.x;
    if (k < 0 || k >= dr.chunk_cap) return;
    const int64_t cg = (int64_t)dr.chunk_base + k;
    if (!ax_emit_active(w, dr, st, k)) return;
    if (mode == 1 && k >= st.k0 + w.nk_full) return;
    if (mode >= 2 && k < st.k0 + w.nk_full) return;
    const bool fused = mode == 2;
    AxChunk& ch = w.chunk[cg];
    const int ne = ch.n_edges;
    if (ne <= 0) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int nhe = ch.n_head_edges, npre = ch.n_pre;
    ax_emit_levels(w, dr, ch, (int)threadIdx.x, (int)blockDim.x);
    for (int t = threadIdx.x; t < nhe; t += blockDim.x) ax_emit_edge(w, dr, st, c, ch, cg, k, t, 0, fused);
    if (threadIdx.x == 0 && npre > 0) {
        const uint8_t* nx = w.zc_nx + dr.zc_base;
        int64_t pos = ch.g_first;
        for (int q = 0; q < npre; ++q) { ax_emit_edge(w, dr, st, c, ch, cg, k, nhe + q, pos, fused); if (q < npre - 1) pos += nx[pos]; }
    }
    if (ch.merge_pos < 0) return;
    const uint64_t* cmask = w.cmask + dr.tile_base;
    const int32_t* crank = w.crank + dr.tile_base;
    const int64_t mp = ch.merge_pos;
    const int64_t t0 = mp / AX_TILE, t1 = ch.q_last / AX_TILE;
    const int64_t r0 = ax_canon_rank(cmask, crank, mp) - (nhe + npre);       // edge number = canonical rank - r0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The tile's mask and running count and the records of all its 64 crossings are fetched in ONE round of loads (the
    // records before the mask says which of them are edges: whole lines, and no second round of dependent loads -- the
    // kernel was bound by that latency); the dense arrays are padded to whole tiles.
    const int32_t* zi = w.zc_idx + dr.zc_base;
    const float* za1 = w.zc_a1 + dr.zc_base;
    const float* za2 = w.zc_a2 + dr.zc_base;
    for (int64_t tt = t0 + warp; tt <= t1; tt += blockDim.x >> 5) {
        const uint64_t cm = cmask[tt];
        const int32_t cr = crank[tt];
        int32_t ri[2]; float r1[2], r2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t pos = tt * AX_TILE + lane + 32 * h;
            ri[h] = zi[pos]; r1[h] = za1[pos]; r2[h] = za2[pos];
        }
        uint64_t m = cm;
        if (tt == t0) m &= ~((1ull << (mp - t0 * AX_TILE)) - 1ull);
        const int64_t base = (int64_t)cr - r0;
        if (base >= ne) break;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bit = lane + 32 * h;
            if ((m >> bit) & 1ull) {
                const int64_t t = base + __popcll(cm & ((1ull << bit) - 1ull));
                if (t < ne) ax_emit_edge_vals(w, dr, st, c, ch, cg, k, (int)t, (int64_t)ri[h], (double)r1[h], (double)r2[h], fused);
            }
        }
    }
}

// ------------------------------------------------------------------ per-drop bookkeeping, one warp per drop
// Same results as ax_verify_item / ax_offsets_item / ax_plan_tones_item / ax_chain_item; the loops over the
// run() iterations of a drop advance 32 iterations per step (ballot / scan), and the chain uses the lanes to
// fetch the handful of table entries each step needs in one round instead of one dependent load after another.
__device__ __forceinline__ int ax_warp_excl_scan(int v, int lane, int* total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    *total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

__global__ void __launch_bounds__(32) k_verify_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || st.chain_end) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int k0 = st.chain_from, k1 = st.n_chunks;
    __syncwarp();
    if (lane == 0) st.chain_dirty = 0;
    for (int kb = k0; kb < k1; kb += 32) {
        const int k = kb + lane;
        int code = 0;                                   // 1 error, 2 start index would become a float, 3 mis-speculated
        if (k < k1) {
            if (ch[k].err) code = 1;
            else if (ch[k].true_last - ch[k].s - 1 <= c.pad) code = 2;
            else if (ch[k].true_last != ch[k].spec_last) code = 3;
        }
        const unsigned ball = __ballot_sync(0xffffffffu, code != 0);
        if (ball) {
            if (lane == __ffs((int)ball) - 1) {         // the first iteration that stops the scan, as the sequential form
                if (code == 1) { ax_raise(st, ch[k].err, k); st.n_chunks = k + 1; st.chain_from = k + 1; st.chain_end = 1; }
                else if (code == 2) {
                    const double sf = (double)ch[k].s + c.fs / (double)c.bitrate;       // AXCTDprocessor.py:331
                    const bool ends = (double)dr.n - sf < 4.0 * c.n_power;
                    if (!ends) ax_raise(st, AXCTD_DROP_FLOAT_INDEX, k + 1);
                    if (ends && w.streaming == 1) { st.n_chunks = k; st.chain_from = k; st.chain_end = 1; }   // the file may end here: a later run decides
                    else { st.n_chunks = k + 1; st.chain_from = k + 1; st.chain_end = 1; }
                } else { st.n_fixups++; st.chain_from = k + 1; st.chain_dirty = 1; w.flags[AX_FLAG_DIRTY] = 1; }
            }
            return;
        }
    }
    if (lane == 0) { st.chain_from = st.n_chunks; st.chain_end = 1; }
}

__global__ void __launch_bounds__(32) k_offsets_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (st.sm_status < 1) { if (lane == 0) { st.nbits_total = 0; st.nedges_total = 0; } return; }
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int k0 = st.k0, k1 = st.n_chunks;
    int64_t nb = 0, ne = 0;
    for (int kb = k0; kb < k1; kb += 32) {
        const int k = kb + lane;
        const int n = (k < k1 && ch[k].n_edges > 0) ? ch[k].n_edges : 0;
        int tot_e, tot_b;
        const int pe = ax_warp_excl_scan(n, lane, &tot_e);
        const int pb = ax_warp_excl_scan(n > 0 ? n - 1 : 0, lane, &tot_b);
        if (k < k1) { ch[k].bit_off = nb + pb; ch[k].edge_off = ne + pe; }
        nb += tot_b; ne += tot_e;
    }
    if (lane == 0) {
        if (ne > dr.edge_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, -1); w.flags[AX_FLAG_CAP] = 1; nb = 0; ne = 0; }
        st.nbits_total = nb; st.nedges_total = ne;
    }
}

__global__ void __launch_bounds__(32) k_plan_tones_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (st.sm_status < 1) return;
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int k0 = st.k0 + 1, k1 = st.n_chunks;
    const int nsm = st.next_sm_chunk;          // (iterations the state machine has been through keep their records: streaming)
    int32_t pc = ch[st.k0].pw_off + ch[st.k0].np;
    bool small = false;
    for (int kb = k0; kb < k1; kb += 32) {
        const int k = kb + lane;
        const int np = k < k1 ? ax_grid_count(ch[k].s, ch[k].e, c) : 0;
        int tot;
        const int pre = ax_warp_excl_scan(np, lane, &tot);
        const bool over = k < k1 && pc + pre + np > dr.pw_cap;
        const unsigned ball = __ballot_sync(0xffffffffu, over);
        const int stop = ball ? __ffs((int)ball) - 1 : 32;      // first lane whose samples do not fit
        if (k < k1 && lane <= stop) {
            AxChunk& q = ch[k];
            q.pw_off = pc + pre; q.np = lane == stop ? 0 : np;
            if (k >= nsm) { q.status = 0; q.n_rows = 0; q.n_hex = 0; q.frame_begin = q.frame_end = 0; q.scale = c.scale0; q.mean7500 = ax_nan(); }
            if (lane < stop && np < 10) small = true;
        }
        if (ball) {
            if (lane == stop) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; st.n_chunks = k; }
            break;
        }
        pc += tot;
    }
    if (__any_sync(0xffffffffu, small) && lane == 0) st.par_levels = 0;
}

// Pulse search of a detection round (ax_sm_item, phase 0, status 0) with the lanes looking at different power samples:
// the first sample whose 400 Hz level reaches the threshold decides where the sequential state machine has anything
// to do; the iterations before it only record "still status 0".  A one-hour recording scanned for its drops
// (segment.scan_batch) has 90 000 power samples and no pulse the scan accepts: one thread took 4 ms per round.
__global__ void __launch_bounds__(32) k_sm_search_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status != 0 || !st.par_levels) return;
    int klo, kend;
    if (!ax_level_range(w, st, 0, &klo, &kend)) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int kfrom = st.next_sm_chunk;
    if (kfrom >= kend) return;
    const double* r400 = w.r400 + dr.pw_base;
    const int i0 = ch[kfrom].pw_off, i1 = ch[kend - 1].pw_off + ch[kend - 1].np;
    int first = i1;
    for (int ib = i0; ib < i1 && first == i1; ib += 32 * 8) {
        int mine = i1;
#pragma unroll
        for (int u = 7; u >= 0; --u) { const int i = ib + 32 * u + lane; if (i < i1 && r400[i] >= c.min_r400) mine = i; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
        first = mine;
    }
    int kf = kend;                                     // iteration that holds the first hit
    if (first < i1) {
        int lo = kfrom, hi = kend - 1;                 // last iteration with pw_off <= first
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (ch[mid].pw_off <= first) lo = mid; else hi = mid - 1; }
        kf = lo;
    }
    for (int k = kfrom + lane; k < kf; k += 32) { ch[k].mean7500 = st.mean7500; ch[k].status = 0; ch[k].profstart = st.profstartind; }
    __syncwarp();
    if (lane == 0 && kf > kfrom) { st.pcount = ch[kf - 1].pw_off + ch[kf - 1].np; st.next_sm_chunk = kf; }
}

// ax_header_item with the window's bits staged in shared memory by the warp (one thread read up to 4 400 bits from
// global memory one byte at a time); the parse itself stays with lane 0.  grid = 2 x drops (header slots)
#define AX_HDR_STAGE 8192
__global__ void __launch_bounds__(32) k_headers_warp(AxWave w) {
    __shared__ uint8_t sb[AX_HDR_STAGE];
    const int d = (int)(blockIdx.x >> 1), slot = (int)(blockIdx.x & 1), lane = threadIdx.x;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    ax_header_reset(st, slot, lane, 32);
    __syncwarp();
    if (st.sm_status < 1 || st.nedges_total == 0) return;
    const uint8_t* B = w.bit + dr.edge_base;
    const int klast = (st.k2 >= 0) ? st.k2 : st.n_chunks - 1;
    for (int k = st.k0; k <= klast && k < st.n_chunks; ++k) {
        int64_t a = 0, n = 0;
        int r = 0;
        if (lane == 0) r = ax_header_window(w, dr, c, st, slot, k, &a, &n);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r < 0) { if (lane == 0) ax_raise(st, -r, k); return; }
        if (r == 0) continue;
        a = __shfl_sync(0xffffffffu, a, 0); n = __shfl_sync(0xffffffffu, n, 0);
        const bool staged = n <= AX_HDR_STAGE;
        if (staged) for (int64_t i = lane; i < n; i += 32) sb[i] = B[a + i];
        __syncwarp();
        int done = 0;
        if (lane == 0) done = ax_header_parse(st, slot, k, staged ? sb : B + a, n) ? 1 : 0;
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done) return;
        __syncwarp();
    }
}

// ax_plan0_item in closed form, a CTA per drop: while the status is 0 the iterations tile the recording
// (start k * chunk, AXCTDprocessor.py:333), so their number and every field follow from k alone.
__global__ void __launch_bounds__(128) k_plan0_block(AxWave w) {
    const int d = blockIdx.x;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    AxChunk* ch = w.chunk + dr.chunk_base;
    __shared__ int s_viol, s_small;
    int kfrom = 0, par0 = 1;
    if (w.streaming) {
        if (st.sm_status >= 1 || st.status != 0) return;
        kfrom = st.n_fixed;
        if (kfrom > 0) par0 = st.par_levels;
    }
    const int64_t CL = c.chunk_len, n = dr.n;
    int64_t K;
    if (w.streaming == 1) K = n > 0 ? (n - 1) / CL : 0;                     // iterations with start + chunk inside the data
    else K = n >= 4 * (int64_t)c.n_power ? (n - 4 * (int64_t)c.n_power) / CL + 1 : 0;     // :295
    bool cap_hit = false;
    if (K > dr.chunk_cap) { K = dr.chunk_cap; cap_hit = true; }
    const int32_t np0 = ax_grid_count(0, CL, c);
    if (threadIdx.x == 0) { s_viol = 0x7fffffff; s_small = 0; }
    __syncthreads();
    for (int64_t k = kfrom + threadIdx.x; k < K; k += blockDim.x) {
        const int64_t s = k * CL;
        int64_t e = s + CL; if (e >= n) e = n - 1;
        const int32_t np = ax_grid_count(s, e, c);
        if (k * (int64_t)np0 + np > dr.pw_cap) atomicMin(&s_viol, (int)k);
    }
    __syncthreads();
    const int64_t Kf = s_viol < K ? s_viol : K;
    for (int64_t k = kfrom + threadIdx.x; k < Kf; k += blockDim.x) {
        const int64_t s = k * CL;
        int64_t e = s + CL; if (e >= n) e = n - 1;
        AxChunk& q = ch[k];
        q.s = s; q.e = e; q.pw_off = (int32_t)(k * np0); q.np = ax_grid_count(s, e, c);
        q.n_edges = 0; q.n_head_edges = 0; q.err = 0; q.status = 0; q.n_rows = 0; q.n_hex = 0;
        q.frame_begin = q.frame_end = 0; q.scale = c.scale0; q.mean7500 = ax_nan();
        q.spec_last = q.true_last = -1; q.g_first = -1; q.q_last = -1; q.bit_off = q.edge_off = 0; q.first_edge = -1;
        if (q.np < 10) atomicOr(&s_small, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_viol < K) { ax_raise(st, AXCTD_DROP_CAPACITY, s_viol); w.flags[AX_FLAG_CAP] = 1; }
        else if (cap_hit) { ax_raise(st, AXCTD_DROP_CAPACITY, (int)K); w.flags[AX_FLAG_CAP] = 1; }
        st.n_fixed = (int32_t)Kf; st.n_chunks = (int32_t)Kf;
        st.par_levels = (par0 && !s_small) ? 1 : 0;
        st.searching = Kf > 0 ? 1 : 0;
        if (w.streaming && st.next_sm_chunk >= Kf) st.searching = 0;
    }
}

// ax_chain_item with the lanes fetching in parallel.  Per run() iteration three rounds of loads: the crossings
// around the predicted end of the chunk (lane i looks at ordinal guess-16+i), the canonical masks of the tiles
// that can hold the stopping crossing, and the two crossing indices that fix the next start.
__global__ void __launch_bounds__(32) k_chain_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    AxState& st = w.st[d];
    if (st.status != 0 || st.sm_status < 1 || st.chain_end) return;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int32_t* zi = w.zc_idx + dr.zc_base;
    const uint8_t* nx = w.zc_nx + dr.zc_base;
    const uint64_t* cmask = w.cmask + dr.tile_base;
    const int64_t M = st.zc_count;
    int k = st.chain_from;
    int64_t s;
    if (k == st.k0) s = ch[k].s;
    else s = ch[k - 1].true_last - 1 - c.pad;
    int64_t entry = -1;
    int n_chunks_out = -1;
    // the crossing that ends an iteration is found through the coarse index zq (crossings below every AX_ZQ-th sample,
    // k_compact_warp): the 64 crossings from zq[(e - 2) / AX_ZQ] - 9 on hold it, its four predecessors and the crossings
    // that fix the next start.  The entries the NEXT iteration can need are fetched in the same round of loads
    // (tq_lo .. tq_lo + 2), so that an iteration costs one dependent memory round trip.
    const int32_t* zq = w.zc_q + dr.zq_base;
    const int64_t nzq = dr.n / AX_ZQ + 2;
    int64_t tq_lo = -1; int32_t tq_val = 0;              // lane i < 3 holds zq[tq_lo + i]
    for (;; ++k) {
        if (w.streaming == 1) { if (s + c.chunk_len >= dr.n) { n_chunks_out = k; break; } }      // not complete yet: a later run takes it
        else if (dr.n - s < 4 * (int64_t)c.n_power) { n_chunks_out = k; break; }
        if (k >= dr.chunk_cap) { if (lane == 0) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; } n_chunks_out = k; break; }
        int64_t e = s + c.chunk_len;
        if (e >= dr.n) e = dr.n - 1;
        if (lane == 0) { ch[k].s = s; ch[k].e = e; ch[k].err = 0; ch[k].n_edges = 0; ch[k].spec_last = -1; }
        if (entry < 0) entry = ax_lower_bound(zi, M, s + c.pad);            // (all lanes: same result)
        // ---- one round of loads: the 32 crossings around the predicted end of the chunk, and (lanes 0..2) the
        // canonical masks of the entry tile and of the two tiles that should hold the stopping crossing
        const int64_t jq = (e - 2) >> AX_ZQ_SHIFT;
        int64_t t0;
        if (jq >= tq_lo && jq < tq_lo + 3 && tq_lo >= 0) t0 = __shfl_sync(0xffffffffu, tq_val, (int)(jq - tq_lo));
        else t0 = zq[jq < nzq ? jq : nzq - 1];
        const int64_t g0 = t0 - 9;
        const int64_t mine = g0 + lane, mine2 = mine + 32;
        const int32_t zmine = (mine >= 0 && mine < M) ? zi[mine] : 0;
        const int32_t zmine2 = (mine2 >= 0 && mine2 < M) ? zi[mine2] : 0;
        {   // entries of the coarse index the next iteration can ask for: its end lies one chunk (less the pad and the
            // few crossings the walk stops short) after this one's
            tq_lo = (e - 2 + c.chunk_len - c.pad - 512 - AX_ZQ) >> AX_ZQ_SHIFT;
            if (tq_lo < 0) tq_lo = 0;
            const int64_t jn = tq_lo + lane;
            tq_val = (lane < 3) ? zq[jn < nzq ? jn : nzq - 1] : 0;
        }
        const int64_t tg = (t0 - 4) / AX_TILE;                                  // likely tile of X
        const int64_t tl = lane == 0 ? entry / AX_TILE : tg + (lane - 1);
        const uint64_t mv = (lane < 3 && tl >= 0) ? cmask[tl] : 0ull;
        int64_t q;
        {
            const bool le = mine < 0 ? true : (mine >= M ? false : (int64_t)zmine <= e - 2);
            const bool le2 = mine2 < 0 ? true : (mine2 >= M ? false : (int64_t)zmine2 <= e - 2);
            const unsigned ball = __ballot_sync(0xffffffffu, le), ball2 = __ballot_sync(0xffffffffu, le2);
            if (ball2 != 0u && ball2 != 0xffffffffu) q = g0 + 32 + (31 - __clz((int)ball2));   // zi is ascending: le is a prefix
            else if (ball2 == 0u && ball != 0u) q = g0 + (31 - __clz((int)ball));
            else q = ax_upper_bound_from(zi, M, e - 2, t0 + 54) - 1;
        }
        if (entry > q) { n_chunks_out = k + 1; break; }
        // ---- where the walk stops (ax_walk_end without the step count)
        const int64_t X = q - 4;
        int64_t pos = entry;
        bool first = true;
        while (pos < X) {
            const int64_t t = pos / AX_TILE, tx = X / AX_TILE;
            uint64_t m_t, m_x, m_x1;
            if (first && tx == tg) {                                           // the prefetched masks are the right ones
                m_t = __shfl_sync(0xffffffffu, mv, 0); m_x = __shfl_sync(0xffffffffu, mv, 1); m_x1 = __shfl_sync(0xffffffffu, mv, 2);
            } else {
                const int64_t tl2 = lane == 0 ? t : tx + (lane - 1);
                const uint64_t mv2 = (lane < 3) ? cmask[tl2] : 0ull;
                m_t = __shfl_sync(0xffffffffu, mv2, 0); m_x = __shfl_sync(0xffffffffu, mv2, 1); m_x1 = __shfl_sync(0xffffffffu, mv2, 2);
            }
            first = false;
            if ((m_t >> (pos - t * AX_TILE)) & 1ull) {                       // on the canonical walk: jump
                uint64_t m = m_x & ~((1ull << (X - tx * AX_TILE)) - 1ull);
                int64_t tt = tx;
                if (!m) { m = m_x1; tt = tx + 1; }
                while (!m) m = cmask[++tt];
                pos = tt * AX_TILE + ax_ctz64(m);
                break;
            }
            pos += nx[pos];
        }
        // ---- the crossing indices that fix the next start: from the window when they are in it
        int64_t zpos, zprev;
        if (pos - 1 >= g0 && pos < g0 + 64 && pos - 1 >= 0 && pos < M) {
            const int ia = (int)(pos - g0), ib = ia - 1;                     // window positions (two registers of 32)
            const int32_t a0 = __shfl_sync(0xffffffffu, zmine, ia & 31), a1 = __shfl_sync(0xffffffffu, zmine2, ia & 31);
            const int32_t b0 = __shfl_sync(0xffffffffu, zmine, ib & 31), b1 = __shfl_sync(0xffffffffu, zmine2, ib & 31);
            zpos = ia < 32 ? a0 : a1;
            zprev = ib < 32 ? b0 : b1;
        } else {
            const int64_t pl = pos - (lane & 1);
            const int32_t zv = (lane < 2 && pl >= 0) ? zi[pl] : 0;
            zpos = __shfl_sync(0xffffffffu, zv, 0); zprev = __shfl_sync(0xffffffffu, zv, 1);
        }
        if (lane == 0) ch[k].spec_last = zpos;
        const int64_t next_ind = zpos - s - 1;                               // demodulate.py:104
        if (next_ind <= c.pad) { n_chunks_out = k + 1; break; }
        s = s + next_ind - c.pad;
        entry = (pos > 0 && zprev >= zpos - 1) ? pos - 1 : pos;
    }
    if (lane == 0) st.n_chunks = n_chunks_out;
}

// ax_frames_chain_item, 32 run() iterations per step: a speculative scan is right when it started where the
// previous non-empty iteration's scan ended; the first one that did not is redone by one lane (rare), after
// which the group is retried from the next iteration.
__global__ void __launch_bounds__(32) k_frames_chain_warp(AxWave w) {
    const int d = blockIdx.x, lane = threadIdx.x;
    const AxDrop& dr = w.drop[d];
    AxState& st = w.st[d];
    if (lane == 0) st.n_frames = 0;
    if (st.status != 0 || st.sm_status < 2 || st.k2 < 0 || st.nedges_total == 0) return;
    AxChunk* ch = w.chunk + dr.chunk_base;
    const int k1 = st.n_chunks;
    int64_t cur = 0;
    int32_t nf = 0;
    int kb = st.k2;
    while (kb < k1) {
        const int k = kb + lane;
        const bool in = k < k1;
        const bool ne = in && ch[k].n_edges > 0;
        const int64_t sfrom = ne ? ch[k].scan_from : 0, send = ne ? ch[k].scan_end : 0;
        const int cnt = ne ? ch[k].scan_cnt : 0;
        const unsigned nem = __ballot_sync(0xffffffffu, ne);
        const unsigned below = nem & ((1u << lane) - 1u);
        const int prev = below ? 31 - __clz((int)below) : -1;
        const int64_t pend = __shfl_sync(0xffffffffu, send, prev < 0 ? 0 : prev);
        const int64_t expect = prev < 0 ? cur : pend;
        const unsigned bad = __ballot_sync(0xffffffffu, ne && sfrom != expect);
        const int nok = bad ? __ffs((int)bad) - 1 : 32;          // lanes [0, nok) are consistent
        int tot;
        const int pre = ax_warp_excl_scan(lane < nok ? cnt : 0, lane, &tot);
        const unsigned over = __ballot_sync(0xffffffffu, in && lane < nok && nf + pre + cnt > dr.frame_cap);
        if (over) {
            if (lane == __ffs((int)over) - 1) { ax_raise(st, AXCTD_DROP_CAPACITY, k); w.flags[AX_FLAG_CAP] = 1; }
            return;
        }
        if (in && lane < nok) { ch[k].frame_begin = nf + pre; ch[k].frame_end = nf + pre + cnt; }
        nf += tot;
        {   // carry: end of the last non-empty consistent iteration
            const unsigned okne = nem & (nok >= 32 ? 0xffffffffu : ((1u << nok) - 1u));
            if (okne) cur = __shfl_sync(0xffffffffu, send, 31 - __clz((int)okne));
        }
        if (nok >= 32) { kb += 32; continue; }
        // iteration kb + nok started from the wrong bit: redo its scan from the true start
        const int kk = kb + nok;
        int err = 0;
        int32_t c2 = 0; int64_t e2 = 0;
        if (lane == 0) {
            err = ax_frames_scan(w, dr, st, ch[kk], kk, cur, nullptr, 0x7fffffff, &c2, &e2);
            if (err) ax_raise(st, err, kk);
            else {
                ch[kk].scan_from = cur; ch[kk].scan_cnt = c2; ch[kk].scan_end = e2; st.n_frame_respec++;
                if (nf + c2 > dr.frame_cap) { ax_raise(st, AXCTD_DROP_CAPACITY, kk); w.flags[AX_FLAG_CAP] = 1; err = 1; }
                else { ch[kk].frame_begin = nf; ch[kk].frame_end = nf + c2; }
            }
        }
        err = __shfl_sync(0xffffffffu, err, 0);
        if (err) return;
        nf += __shfl_sync(0xffffffffu, c2, 0);
        cur = __shfl_sync(0xffffffffu, e2, 0);
        kb = kk + 1;
    }
    if (lane == 0) st.n_frames = nf;
}

// ------------------------------------------------------------------ scale calibration (CTA per drop)
// ax_scale_item with the histogram filled by the whole CTA (shared-memory atomics).
__global__ void __launch_bounds__(128) k_scale_block(AxWave w) {
    const int d = blockIdx.x;
    const AxDrop& dr = w.drop[d];
    const AxCfg& c = w.cfg[dr.cfg];
    AxState& st = w.st[d];
    __shared__ int s_hist[512];
    __shared__ int s_found, s_k;
    __shared__ int64_t s_a, s_hi;
    const int nbins = c.n_hist_edges - 1;
    if (ax_scale_is_final(w, st)) { ax_scale_spread(w, d, threadIdx.x, blockDim.x); return; }      // (uniform over the CTA)
    if (threadIdx.x == 0) {
        ax_scale_reset(w, d);
        int k = -1; int64_t a = 0, hi = 0;
        int found = 0;
        if (st.sm_status >= 1 && st.nedges_total != 0) {
            found = ax_scale_find(w, d, &k, &a, &hi);
            if (found < 0) ax_raise(st, -found, k);
            else if (found && nbins > 512) { ax_raise(st, AXCTD_DROP_CAPACITY, k); found = -1; }
        } else found = -1;
        s_found = found; s_k = k; s_a = a; s_hi = hi;
    }
    for (int q = threadIdx.x; q < 512; q += blockDim.x) s_hist[q] = 0;
    __syncthreads();
    if (s_found < 0) return;
    if (s_found) {
        const double* a1 = w.a1 + dr.edge_base; const double* a2 = w.a2 + dr.edge_base;
        for (int64_t jj = s_a + threadIdx.x; jj < s_hi; jj += blockDim.x) {
            const int bin = ax_scale_bin(c, ax_div(ax_mul(a2[jj], c.scale0), a1[jj]));
            if (bin >= 0) atomicAdd(&s_hist[bin], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double thr;
            if (!ax_scale_threshold(c, s_hist, s_hi > s_a ? s_hi - s_a : 0, &thr)) { ax_raise(st, AXCTD_DROP_SCALE_EMPTY, s_k); s_found = -1; }
            else { st.scale = ax_div(c.scale0, thr); st.k1 = s_k; st.header_read[0] = 1; st.header_chunk[0] = s_k; }
        }
        __syncthreads();
        if (s_found < 0) return;
    }
    ax_scale_spread(w, d, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------ spike filter (warp per run() iteration)
// ax_qc_item with the percentiles taken by rank counting across the warp (up to 64 kept rows per
// iteration, two per lane); larger iterations fall back to the one-thread form.
// ranks (0-based, ties broken by position: slot 0 of lane j is entry j, slot 1 is entry j + 32) of this lane's two entries
__device__ __forceinline__ void ax_warp_ranks(double v0, double v1, bool has0, bool has1, int lane, int* r0_out, int* r1_out) {
    int r0 = 0, r1 = 0;
    for (int j = 0; j < 32; ++j) {
        const double u0 = __shfl_sync(0xffffffffu, v0, j), u1 = __shfl_sync(0xffffffffu, v1, j);
        const bool h0 = __shfl_sync(0xffffffffu, (int)has0, j), h1 = __shfl_sync(0xffffffffu, (int)has1, j);
        if (h0) { r0 += (u0 < v0) || (u0 == v0 && j < lane); r1 += (u0 < v1) || (u0 == v1); }
        if (h1) { r0 += (u1 < v0); r1 += (u1 < v1) || (u1 == v1 && j < lane); }
    }
    *r0_out = r0; *r1_out = r1;
}
__device__ __forceinline__ double ax_warp_kth(double v0, double v1, bool has0, bool has1, int r0, int r1, int kth) {
    const unsigned b0 = __ballot_sync(0xffffffffu, has0 && r0 == kth), b1 = __ballot_sync(0xffffffffu, has1 && r1 == kth);
    if (b0) return __shfl_sync(0xffffffffu, v0, __ffs((int)b0) - 1);
    return __shfl_sync(0xffffffffu, v1, __ffs((int)b1) - 1);
}
__device__ __forceinline__ double ax_warp_percentile(double v0, double v1, bool has0, bool has1, int r0, int r1, int n, double q) {
    const double virt = ax_mul((double)(n - 1), q);                  // ax_percentile_sorted
    int prev = (int)floor(virt), next = prev + 1;
    if (virt >= (double)(n - 1)) { prev = n - 1; next = n - 1; }
    if (virt < 0) { prev = 0; next = 0; }
    const double gamma = ax_sub(virt, floor(virt));
    const double a = ax_warp_kth(v0, v1, has0, has1, r0, r1, prev), b = ax_warp_kth(v0, v1, has0, has1, r0, r1, next);
    const double diff = ax_sub(b, a);
    double r = ax_add(a, ax_mul(diff, gamma));
    if (gamma >= 0.5) r = ax_sub(b, ax_mul(diff, ax_sub(1.0, gamma)));
    return r;
}
__global__ void __launch_bounds__(128) k_qc_warp(AxWave w, double* scratch) {
    const int64_t cg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (cg >= gridDim.x * (int64_t)(blockDim.x >> 5)) return;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::chunk_base, cg);
    const AxDrop& dr = w.drop[d];
    const AxState& st = w.st[d];
    const int k = (int)(cg - dr.chunk_base);
    if (k >= dr.chunk_cap || st.status != 0 || st.k2 < 0 || k < st.k2 || k >= st.n_chunks) return;
    AxChunk& ch = w.chunk[cg];
    const int nfr = ch.frame_end - ch.frame_begin;
    if (nfr > 64) { if (lane == 0) ax_qc_item(w, cg, scratch); return; }
    if (lane == 0) { ch.n_rows = 0; ch.n_hex = 0; }
    if (nfr <= 0) return;
    axctd_frame* fr = w.frame + dr.frame_base + ch.frame_begin;
    const int i0 = lane, i1 = lane + 32;
    const bool has0 = i0 < nfr && fr[i0].keep, has1 = i1 < nfr && fr[i1].keep;
    const double T0 = has0 ? fr[i0].temperature : 0.0, T1 = has1 ? fr[i1].temperature : 0.0;
    const double S0 = has0 ? fr[i0].salinity : 0.0, S1 = has1 ? fr[i1].salinity : 0.0;
    const int n = __popc(__ballot_sync(0xffffffffu, has0)) + __popc(__ballot_sync(0xffffffffu, has1));
    if (n == 0) return;
    const bool tnan = __any_sync(0xffffffffu, (has0 && isnan(T0)) || (has1 && isnan(T1)));
    const bool snan = __any_sync(0xffffffffu, (has0 && isnan(S0)) || (has1 && isnan(S1)));
    double Tlo = ax_nan(), Thi = ax_nan(), Slo = ax_nan(), Shi = ax_nan();
    int r0, r1;
    if (!tnan) {
        ax_warp_ranks(T0, T1, has0, has1, lane, &r0, &r1);
        const double m = ax_warp_percentile(T0, T1, has0, has1, r0, r1, n, 0.5);
        Tlo = ax_sub(m, ax_mul(10.0, ax_sub(m, ax_warp_percentile(T0, T1, has0, has1, r0, r1, n, 0.15))));
        Thi = ax_add(m, ax_mul(10.0, ax_sub(ax_warp_percentile(T0, T1, has0, has1, r0, r1, n, 0.85), m)));
    }
    if (!snan) {
        ax_warp_ranks(S0, S1, has0, has1, lane, &r0, &r1);
        const double m = ax_warp_percentile(S0, S1, has0, has1, r0, r1, n, 0.5);
        Slo = ax_sub(m, ax_mul(10.0, ax_sub(m, ax_warp_percentile(S0, S1, has0, has1, r0, r1, n, 0.15))));
        Shi = ax_add(m, ax_mul(10.0, ax_sub(ax_warp_percentile(S0, S1, has0, has1, r0, r1, n, 0.85), m)));
    }
    const bool k0 = has0 && !(T0 < Tlo || T0 > Thi || S0 < Slo || S0 > Shi);
    const bool k1 = has1 && !(T1 < Tlo || T1 > Thi || S1 < Slo || S1 > Shi);
    if (has0 && !k0) fr[i0].keep = 0;
    if (has1 && !k1) fr[i1].keep = 0;
    const int rows = __popc(__ballot_sync(0xffffffffu, k0)) + __popc(__ballot_sync(0xffffffffu, k1));
    if (rows > 0) {                                                    // :611-612
        if (i0 < nfr) fr[i0].hex_returned = 1;
        if (i1 < nfr) fr[i1].hex_returned = 1;
    }
    if (lane == 0) { ch.n_rows = rows; ch.n_hex = rows > 0 ? nfr : 0; }
}

// ------------------------------------------------------------------ dense crossing arrays
// ax_compact_item with a warp per segment (coalesced) together with the walk steps (ax_nx_item).
__global__ void __launch_bounds__(256) k_compact_warp(AxWave w) {
    const int64_t seg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (seg >= w.nseg_total) return;
    const int lane = threadIdx.x & 31;
    const int d = w.seg_drop[seg];
    const AxDrop& dr = w.drop[d];
    const int64_t M = w.st[d].zc_count;
    if (M == 0) return;
    const AxCfg& c = w.cfg[dr.cfg];
    const int64_t src = seg * (int64_t)w.seg_cap, dst = dr.zc_base + w.seg_off[seg] + w.blk_sum[seg / 128];
    const int cnt = w.seg_cnt[seg];
    auto fill_zq = [&]() {   // coarse index (k_chain_warp): zq[j] = number of crossings of the drop with index below AX_ZQ * j.  A segment
        // owns the boundaries inside its own sample range (its records are exactly the crossings in that range, so the
        // entry is its dense offset plus the number of its records below the boundary); the last one also closes the table
        const int64_t jseg = seg - dr.seg_base;
        if (jseg < dr.nseg) {
            const int64_t lo_s = jseg * (int64_t)w.seg_len;
            const int64_t j0 = (lo_s + AX_ZQ - 1) >> AX_ZQ_SHIFT;
            const int64_t j1 = jseg == dr.nseg - 1 ? dr.n / AX_ZQ + 2 : (lo_s + w.seg_len + AX_ZQ - 1) >> AX_ZQ_SHIFT;    // exclusive
            int32_t* zq = w.zc_q + dr.zq_base;
            const int32_t base = (int32_t)(dst - dr.zc_base);
            for (int64_t j = j0 + lane; j < j1; j += 32) {
                const int64_t bnd = j << AX_ZQ_SHIFT;
                int lo = 0, hi = cnt;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int64_t)w.rec_idx[src + mid] < bnd) lo = mid + 1; else hi = mid; }
                zq[j] = base + lo;
            }
        }
    };
    if (cnt == 0) { fill_zq(); return; }
    // the walk step of every crossing (ax_nx_item) is formed here as well: it needs the next four crossings, which
    // are this segment's own records or, for its last four, the first records of the segments that follow.
    // 128 records per warp and round, all their loads issued before the first use (the kernel is bound by memory
    // latency: one round of 32 records at a time left two thirds of the bandwidth unused).
    const int64_t pos0 = dst - dr.zc_base;
    const int64_t br2 = 2 * (int64_t)c.bitrate;
    // (2^20 * br2 stays below 2^31 and beyond 2 * fs2: every real rate)
    const bool narrow = br2 > 0 && br2 < 2048 && c.fs2 > 0 && c.fs2 < (1 << 28) && (br2 << 20) > 2 * c.fs2;
    const int32_t br2n = (int32_t)br2, fs2n = (int32_t)c.fs2;
    int32_t la[4] = {0, 0, 0, 0};
    bool have_la = false;
    for (int b0 = 0; b0 < cnt; b0 += 128) {
        int32_t z0[4]; float v1[4], v2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = b0 + 32 * u + lane;
            z0[u] = 0; v1[u] = 0.f; v2[u] = 0.f;
            if (q < cnt) { z0[u] = w.rec_idx[src + q]; v1[u] = w.rec_a1[src + q]; v2[u] = w.rec_a2[src + q]; }
        }
        if (!have_la && b0 + 128 + 4 >= cnt) {           // the round(s) that hold the segment's last four records
            int got = 0;
            const int64_t seg_end = (int64_t)dr.seg_base + dr.nseg;
            for (int64_t s2 = seg + 1; got < 4 && s2 < seg_end; ++s2) {
                const int c2 = w.seg_cnt[s2];
                for (int j = 0; j < c2 && got < 4; ++j) la[got++] = w.rec_idx[s2 * (int64_t)w.seg_cap + j];
            }
            have_la = true;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = b0 + 32 * u + lane;
            if (q >= cnt) continue;
            w.zc_idx[dst + q] = z0[u];
            w.zc_a1[dst + q] = v1[u];
            w.zc_a2[dst + q] = v2[u];
            uint8_t nx = 0;
            if (pos0 + q + 4 < M) {                      // as ax_next: nearest to one bit period, first on ties
                int bj = 0;
                if (narrow) {
                    // 32-bit form: distances are capped at 2^20 samples, far beyond the bit period, where |d * br2 - fs2|
                    // only grows with d (a capped candidate never beats an earlier one, capped or not: same choice)
                    int32_t best = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int qq = q + 1 + j;
                        const int32_t z = qq < cnt ? w.rec_idx[src + qq] : la[qq - cnt];
                        const int32_t dd = abs(min(z - z0[u], 1 << 20) * br2n - fs2n);
                        if (j == 0 || dd < best) { best = dd; bj = j; }
                    }
                } else {
                    int64_t best = 0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int qq = q + 1 + j;
                        const int32_t z = qq < cnt ? w.rec_idx[src + qq] : la[qq - cnt];
                        int64_t dd = ((int64_t)z - z0[u]) * br2 - c.fs2;
                        if (dd < 0) dd = -dd;
                        if (j == 0 || dd < best) { best = dd; bj = j; }
                    }
                }
                nx = (uint8_t)(1 + bj);
            }
            w.zc_nx[dst + q] = nx;
        }
    }
    fill_zq();                                           // (after the copy: the records are in cache by now)
}

// ------------------------------------------------------------------ walk tiles from registers
// ax_tiles_item with the tile's 64 walk steps fetched up front (four 16-byte loads) and packed four bits each into
// four 64-bit registers: the four walks through the tile then run without a dependent memory load per step (the
// one-load-per-step form was latency-bound: 0.44 ms for 0.15 GB of traffic).
__global__ void __launch_bounds__(128) k_tiles_reg(int64_t n, AxWave w) {
    const int64_t tg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tg >= n) return;
    const int d = ax_find_owner(w.drop, w.n_drops, &AxDrop::tile_base, tg);
    const AxDrop& dr = w.drop[d];
    const int64_t t = tg - dr.tile_base;
    if (t >= dr.tile_cap) return;
    const int64_t M = w.st[d].zc_count;
    const int64_t first = t * AX_TILE;
    if (first >= M) return;
    const uint4* src = reinterpret_cast<const uint4*>(w.zc_nx + dr.zc_base + first);      // zc_base and first are multiples of 64
    uint64_t pk[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const uint4 q = src[v];
        const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
        uint64_t acc = 0;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t x = wd[h];
            const uint32_t nib = (x & 0xFu) | ((x >> 4) & 0xF0u) | ((x >> 8) & 0xF00u) | ((x >> 12) & 0xF000u);
            acc |= (uint64_t)nib << (16 * h);
        }
        pk[v] = acc;                                     // steps of crossings 16 v .. 16 v + 15, four bits each
    }
    uint32_t map = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        uint64_t mask = 0;
        uint32_t ex = 0xFF;
        int c = o;                                       // tile-relative
        while (first + c < M) {
            if (c >= AX_TILE) { ex = (uint32_t)(c - AX_TILE); break; }
            mask |= 1ull << c;
            const uint64_t word = c < 16 ? pk[0] : c < 32 ? pk[1] : c < 48 ? pk[2] : pk[3];
            const int step = (int)((word >> ((c & 15) * 4)) & 0xFull);
            if (!step) break;
            c += step;
        }
        w.tile_mask[tg * 4 + o] = mask;
        map |= ex << (8 * o);
    }
    w.tile_map[tg] = map;
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

