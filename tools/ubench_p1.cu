// Where the cascade phase of k_demod_fused spends its cycles: the numerators-first low-pass cascade of the production kernel
// (per sample: 4 IDP.2A on packed int16 words, I2F, 7 DFMA, F2F, sign funnel, guard minimum, ring store) as a stand-alone
// loop over rows staged in shared memory, 2 CTAs x 4 warps per SM as in the engine, with pieces removed one at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_p1 tools/ubench_p1.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Coef { double a1[3], a2[3], k0, k1_64; };

// VAR bit 0: no sign funnel / guard minimum; bit 1: no F2F / ring store (y summed in double instead);
// bit 2: u read from shared memory as int32 (no IDP.2A); bit 3: u by running sums (six additions) instead of IDP.2A;
// bit 4: double -> float by bit operations (truncating, flush below the float range) instead of F2F; bit 5: int -> double by
// the 2^52 + 2^31 bit pattern and one DADD instead of I2F
template <int VAR>
__global__ void __launch_bounds__(128, 2) k_p1(double* out, int rows, const __grid_constant__ Coef cc) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int16_t* stage = reinterpret_cast<int16_t*>(smem) + warp * (32 * 72);                   // [lane][72]
    float4* ring = reinterpret_cast<float4*>(smem + 4 * 32 * 72 * 2) + warp * (32 * 32);      // [quad][lane]
    for (int i = lane; i < 32 * 72; i += 32) stage[i] = (int16_t)((i * 2654435761u) >> 17);
    __syncwarp();
    double z0[3] = {0, 0, 0}, z1[3] = {0, 0, 0};
    int3 hist = make_int3(0, 0, 0);
    int r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0;      // running sums (VAR bit 3)
    unsigned long long S = 0;
    float minabs = 1e30f;
    double ysum = 0.0;
    const int4* rp = reinterpret_cast<const int4*>(stage + lane * 72);
    float4* yo = ring + lane;
    for (int t = 0; t < rows; ++t) {
        double pipe[3];
        int4 q = rp[0];
        int4 qp = make_int4(0, hist.x, hist.y, hist.z);
        unsigned sb = 0u;
        float yf[4];
#pragma unroll
        for (int n = 0; n < 64 + 2; ++n) {
#pragma unroll
            for (int s = 2; s >= 0; --s) {
                const int m = n - s;
                if (m >= 0 && m < 64) {
                    double tt;
                    if (s == 0) {
                        if ((m & 7) == 0 && m > 0) { qp = q; q = rp[m >> 3]; }
                        const int c = (m >> 1) & 3;
                        const int cw[4] = {q.x, q.y, q.z, q.w}, pw[4] = {qp.x, qp.y, qp.z, qp.w};
                        int u;
                        if (VAR & 4) u = cw[c] + m;
                        else if (VAR & 8) {
                            const int x = (m & 1) ? (cw[c] >> 16) : (int)(short)(cw[c] & 0xffff);
                            const int s1 = x + r1, s2 = s1 + r2, s3 = s2 + r3, s4 = s3 + r4, s5 = s4 + r5;
                            u = s5 + r6;
                            r1 = x; r2 = s1; r3 = s2; r4 = s3; r5 = s4; r6 = s5;
                        } else {
                            const int w0 = cw[c];
                            const int w1 = c >= 1 ? cw[c - 1] : pw[c + 3];
                            const int w2 = c >= 2 ? cw[c - 2] : pw[c + 2];
                            const int w3 = c >= 3 ? cw[c - 3] : pw[c + 1];
                            if ((m & 1) == 0) u = __dp2a_lo(w3, 0x0601, __dp2a_lo(w2, 0x140F, __dp2a_lo(w1, 0x060F, __dp2a_lo(w0, 0x0001, 0))));
                            else u = __dp2a_lo(w3, 0x0100, __dp2a_lo(w2, 0x0F06, __dp2a_lo(w1, 0x0F14, __dp2a_lo(w0, 0x0106, 0))));
                        }
                        double ud;
                        if (VAR & 32) ud = __hiloint2double(0x43300000, u ^ 0x80000000) - 4503601774854144.0;      // 2^52 + 2^31
                        else ud = (double)u;
                        tt = fma(ud, cc.k0, cc.k1_64);
                    } else tt = pipe[s];
                    const double y = fma(cc.a1[s], z0[s], fma(cc.a2[s], z1[s], tt));
                    z1[s] = z0[s]; z0[s] = y;
                    if (s < 2) pipe[s + 1] = y;
                    else {
                        if (!(VAR & 1)) sb = __funnelshift_l((unsigned)__double2hiint(y), sb, 1);
                        if (VAR & 2) ysum += y;
                        else {
                            float f;
                            if (VAR & 16) {
                                const unsigned hi = (unsigned)__double2hiint(y), lo = (unsigned)__double2loint(y);
                                const unsigned fs = __funnelshift_l(lo, hi, 3);
                                const unsigned fb = ((fs & 0x7fffffffu) ^ 0x40000000u) | (hi & 0x80000000u);
                                f = __uint_as_float((hi & 0x7ff00000u) < 0x38100000u ? (hi & 0x80000000u) : fb);
                            } else f = (float)y;
                            yf[m & 3] = f;
                            if (!(VAR & 1)) minabs = fminf(minabs, fabsf(f));
                            if ((m & 3) == 3) yo[32 * ((m >> 2) + 16 * (t & 1))] = make_float4(yf[0], yf[1], yf[2], yf[3]);
                        }
                        if (!(VAR & 1) && (m & 31) == 31) { S ^= (unsigned long long)__brev(sb) << (m & 32); sb = 0u; }
                    }
                }
            }
        }
        const int4 ql = rp[7]; hist = make_int3(ql.y, ql.z, ql.w);
        __syncwarp();
    }
    if (ysum == 1.2345 || minabs == 1.2345f || S == 0x1234ull || z0[2] == 1.2345) out[threadIdx.x] = ysum + (double)S;
}

template <int VAR>
static void run(const char* what) {
    double* d; cudaMalloc(&d, 4096);
    Coef h;
    const double A1[3] = {1.6926643005998814, 1.7591969461508574, 1.8877140066455618}, A2[3] = {-0.71770845316494558, -0.78522549575504252, -0.91564405607407828};
    for (int s = 0; s < 3; ++s) { h.a1[s] = A1[s]; h.a2[s] = A2[s]; }
    h.k0 = 2.8447757653552447e-07 / 20000.0; h.k1_64 = 64e-9;
    const int rows = 400;
    const size_t smem = 4 * 32 * 72 * 2 + 4 * 32 * 32 * 16 + 40000;      // staging + rings + padding: two CTAs per SM as in the engine
    cudaFuncSetAttribute(k_p1<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_p1<VAR><<<148 * 2, 128, smem>>>(d, rows, h);
    cudaEventRecord(e0);
    k_p1<VAR><<<148 * 2, 128, smem>>>(d, rows, h);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_p1<VAR>);
    printf("%-58s %5.1f SMSP-cycles per warp-sample (2 warps per scheduler, %d registers) %s\n", what, ms * 1e-3 * 1.965e9 / (2.0 * rows * 64), fa.numRegs,
           cudaGetLastError() == cudaSuccess ? "" : "LAUNCH ERROR");
    cudaFree(d);
}

int main() {
    run<0>("full cascade phase");
    run<1>("  without sign funnel and guard minimum");
    run<2>("  without F2F and ring store");
    run<3>("  without both");
    run<4>("  u from a register (no IDP.2A)");
    run<7>("  DFMA + I2F only");
    run<8>("  u by six running additions instead of four IDP.2A");
    run<16>("  double -> float by bit operations instead of F2F");
    run<32>("  int -> double by bit pattern + DADD instead of I2F");
    run<48>("  both conversions off the XU pipe");
    return 0;
}
