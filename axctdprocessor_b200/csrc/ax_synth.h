// ax_synth.h -- bench / test tooling: device-side twin of synth.py.
//
// Generates the synthetic AXCTD drops of BASELINE.json's configs directly in
// HBM (a 1024-drop batch is ~68 GB and cannot be staged through the host).
// Integer arithmetic plus uncontracted IEEE double operations only, so the
// PCM is bit-identical to synth.generate_drop() for the same seed.
#pragma once
#include "ax_types.h"

struct AxSynth {
    int64_t n_total, n0, tone_start;
    int64_t fs;
    uint64_t key1, key2;
    double nscale, gain, tone_amp;
    double sin_coef[9];
    const uint8_t* bits;     // per bit slot
    const uint8_t* gate;
    const uint8_t* par;      // parity of the number of ones before the slot
    int64_t nslots;
    int16_t* out;
};

AX_HD uint64_t ax_mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
AX_HD uint64_t ax_hash64(uint64_t key, uint64_t idx) { return ax_mix64(ax_mix64(idx * 0x9E3779B97F4A7C15ull + key) + key); }

AX_HD double ax_sin_turns(int64_t num, int64_t den, const double* cf) {
    int64_t x4 = 4 * num;
    const bool neg = x4 > 2 * den;
    if (neg) x4 = 4 * den - x4;
    if (x4 > den) x4 = 2 * den - x4;
    const double y = ax_div((double)x4, (double)den);
    const double y2 = ax_mul(y, y);
    double acc = cf[8];
    for (int k = 7; k >= 0; --k) acc = ax_add(ax_mul(acc, y2), cf[k]);
    const double s = ax_mul(acc, y);
    return neg ? -s : s;
}

AX_HDN inline void ax_synth_item(const AxSynth& g, int64_t n) {
    const uint64_t h[2] = {ax_hash64(g.key1, (uint64_t)n), ax_hash64(g.key2, (uint64_t)n)};
    int64_t acc = 0;
    for (int q = 0; q < 2; ++q)
        for (int sh = 0; sh < 64; sh += 16) acc += (int64_t)((h[q] >> sh) & 0xFFFFull);
    double x = ax_mul(ax_sub((double)acc, 262140.0), g.nscale);
    const int64_t rel = n - g.n0;
    if (rel >= 0) {
        const int64_t numer = rel * 800;
        const int64_t b = numer / g.fs, rem = numer - b * g.fs;
        if (b < g.nslots && g.gate[b]) {
            int64_t q = (int64_t)g.par[b] * g.fs + (g.bits[b] == 1 ? 1 : 2) * rem;
            q %= 2 * g.fs;
            x = ax_add(x, ax_sin_turns(q, 2 * g.fs, g.sin_coef));
        }
    }
    const int64_t relt = n - g.tone_start;
    if (relt >= 0) {
        const int64_t qt = (relt * 7500) % g.fs;
        x = ax_add(x, ax_mul(g.tone_amp, ax_sin_turns(qt, g.fs, g.sin_coef)));
    }
    double v = rint(ax_mul(x, g.gain));
    if (v > 32767.0) v = 32767.0;
    if (v < -32767.0) v = -32767.0;
    g.out[n] = (int16_t)v;
}
