/*
 * axctd.h -- C ABI of the B200-native AXCTD demodulation / decoding engine.
 *
 * The reference (cdens/AXCTDprocessor) is pure Python and has no FFI of its
 * own; the drop-in boundary is its Python API (class AXCTD_Processor,
 * reference AXCTDprocessor.py:80-627, driven by processAXCTD.py:126-183).
 * This header is the C boundary underneath the Python mirror of that API
 * (axctdprocessor_b200/AXCTDprocessor.py): plain pointers and sizes, caller
 * owns all host memory, every call returns an int status and never throws.
 * One engine per GPU, one host thread per engine.
 *
 * Each entry point cites the reference code it replaces.
 */
#ifndef AXCTD_H
#define AXCTD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AXCTD_ABI_VERSION 4
#define AXCTD_MAX_SECTIONS 6

/* ---- return codes of API calls ---------------------------------------- */
#define AXCTD_OK                 0
#define AXCTD_ERR_CUDA           1   /* CUDA runtime error, see axctd_last_error */
#define AXCTD_ERR_ARG            2
#define AXCTD_ERR_CAPACITY       3   /* an internal capacity was exceeded */
#define AXCTD_ERR_STATE          4

/* ---- per-drop status (axctd_drop_summary.status) ---------------------- */
/* Non-zero values >= 16 mirror the exception the reference would raise.   */
#define AXCTD_DROP_OK                  0
#define AXCTD_DROP_NO_CROSSING        16  /* IndexError, demodulate.py:85 */
#define AXCTD_DROP_FLOAT_INDEX        17  /* TypeError after AXCTDprocessor.py:331 */
#define AXCTD_DROP_SHORT_WINDOW       18  /* ValueError (broadcast), demodulate.py:101 */
#define AXCTD_DROP_HEADER_VALUE       19  /* ValueError, parse.py:278 */
#define AXCTD_DROP_SCALE_EMPTY        20  /* ValueError (min of empty), demodulate.py:149 */
#define AXCTD_DROP_TRIM_INDEX         21  /* IndexError, AXCTDprocessor.py:546 / :462 */
#define AXCTD_DROP_CAPACITY           32  /* engine capacity exceeded (not a reference error) */
#define AXCTD_DROP_UNCERTAIN          33  /* a sign/threshold decision fell inside the guard band */
#define AXCTD_DROP_CHAIN_DIVERGED     34  /* chunk-chain fix-up did not converge */

/*
 * Settings of one "rate class" (everything AXCTD_Processor derives from f_s
 * and its settings dict): reference AXCTDprocessor.py:117-182 (constants),
 * :187-208 (defaults), :212-262 (derived tables and the Butterworth SOS).
 * Tables are computed by the host with numpy/scipy exactly as the reference
 * does and passed as data.
 */
typedef struct axctd_config_desc {
    double fs;                  /* effective sampling rate f_s (after any /2 decimation) */
    int32_t n_power;            /* int(f_s/10)                      :153 */
    int32_t d_pcm;              /* int(round(f_s/25))               :155 */
    int32_t npcm;               /* N - 2*bit_inset                  :170-171 */
    int32_t chunk_len;          /* int(refreshrate*f_s)             :222 */
    int32_t pad;                /* demod_Npad = 100                 :158 */
    int32_t bit_inset;          /* 1                                :165 */
    int32_t bitrate;            /* 800                              :164 */
    int32_t n_sections;         /* rows of sos (3 lowpass, 6 bandpass) :254-257 */
    double sos[AXCTD_MAX_SECTIONS][6]; /* scipy.signal.butter(..., output='sos') */
    double max_pole_radius;     /* largest |pole| of sos; sizes the warm-up overlap */
    /* cos/sin tables computed by the host with numpy exactly as the reference does: */
    const double* bit_cs;       /* [bit_cs_len][4] = cos(trig1), sin(trig1), cos(trig2), sin(trig2);
                                   trig = 2*pi*n/f_s*f (:245-246), extended to n < bit_cs_len (> npcm) */
    int32_t bit_cs_len;
    int32_t reserved0;
    const double* tone_cs;      /* [n_power][6] = cos,sin of theta400, theta7500, thetadead (:260-262) */
    double min_r400;            /* settings['minr400']              :225 */
    double min_dr7500;          /* settings['mindr7500']            :227 */
    double trigger_from_s;      /* triggerrange[0] (30)             :250 */
    double trigger_to_s;        /* triggerrange[1] (-1)             :250 */
    double high_bit_scale0;     /* 1.5                              :161 */
    double zcoeff[4], tcoeff[4], ccoeff[4];   /* defaults           :203-205 */
    double tlims[2], slims[2];  /*                                  :207-208 */
    const double* temp_lut;     /* [lut_len] parse.read_temp_LUT    parse.py:139 */
    int32_t lut_len;
    const double* hist_edges;   /* [n_hist_edges] np.arange(0,3,0.01)  demodulate.py:130 */
    const double* hist_centers; /* [n_hist_edges-1]                    demodulate.py:132 */
    int32_t n_hist_edges;
    /* /2 decimation of recordings above 50 kHz (AXCTDprocessor.py:60-62): scipy.signal.decimate(pcm, 2) =
     * sosfiltfilt(cheby1(8, 0.05, 0.4, output='sos'), pcm)[::2] with odd padding; fs above is then f_s/2. */
    int32_t decimate;           /* 0 / 1: none, 2: the batch is given the raw recording and halves it on the device,
                                   3: the batch is given the normalised double-precision signal (axctd_batch_upload_f64) */
    int32_t decim_sections;     /* rows of decim_sos (4) */
    int32_t decim_padlen;       /* sosfiltfilt's default padlen (27) */
    double  decim_sos[AXCTD_MAX_SECTIONS][6];
    double  decim_zi[AXCTD_MAX_SECTIONS][2];   /* scipy.signal.sosfilt_zi(decim_sos) */
    double  decim_pole_radius;  /* largest |pole| of decim_sos; sizes the warm-up overlap of the segmented passes */
} axctd_config_desc;

/* Result header of one drop: the scalar attributes processAXCTD.py:149-168 reads. */
typedef struct axctd_drop_summary {
    int32_t status;             /* AXCTD_DROP_* */
    int32_t status_chunk;       /* chunk index where status was raised, or -1 */
    int64_t numpoints;          /* len(audiostream)                 :91 */
    double  f_s;
    int64_t firstpulse400;      /* -1 if never found                :140,378 */
    int64_t profstartind;       /* -1 if never triggered            :141,401 */
    double  firstpointtime;
    double  mean7500pwr;        /*                                  :393 */
    double  high_bit_scale;     /* after header 1                   :467 */
    int32_t n_chunks;           /* loop iterations of run()         :283 */
    int32_t first_demod_chunk;  /* -1 if none */
    int32_t profile_chunk;      /* iteration where status became 2, -1 if none */
    int32_t header_read[3];     /* header1_read..header3_read       :122-124 */
    int32_t header_chunk[3];
    int64_t n_bits;             /* total demodulated bits (all chunks) */
    int64_t n_edges;            /* total bit edges (n_bits + demod chunks) */
    int64_t n_power;            /* len(power_inds) */
    int64_t n_frames;           /* CRC-valid profile frames found */
    int64_t n_rows;             /* rows surviving QC + spike filter */
    int64_t n_hex;              /* hexframes returned to the caller (:612 quirk) */
    int64_t n_crossings;        /* zero crossings of the continuous filter pass */
    int32_t n_uncertain;        /* filter outputs inside the guard band whose sign the exact recomputation did not confirm */
    int32_t n_chain_fixups;     /* mis-speculated chunks repaired */
    int32_t n_guard_hits;       /* filter outputs inside the guard band seen by the fast passes (anywhere in the recording) */
    int32_t n_guard_confirmed;  /* those inside demodulated iterations whose sign scipy-order arithmetic confirmed */
    int64_t pcm_sum;            /* integer sum and max|x| of the int16 input (:55-56) */
    int32_t pcm_ampl;
    int32_t n_recheck;          /* fp32 bit windows re-evaluated in double precision (decision within tolerance) */
    float   win32_max_rel_err;  /* largest relative |fp32 - fp64| window magnitude difference seen at those */
    int32_t n_frame_respec;     /* iterations whose speculative frame scan started from the wrong bit and was redone */
    /* header frames as decoded by parse_header (parse.py:197-285), slot 0 = header 2, 1 = header 3 */
    uint16_t frame_data[2][72];
    uint8_t  counter_found[2][72];
    int32_t  header_parsed[2];  /* parse_header ran for that slot */
    /* merged metadata after AXCTDprocessor.py:505-535 */
    double  zcoeff[4], tcoeff[4], ccoeff[4];      /* metadata['?coeff'] */
    int32_t zcoeff_valid[4], tcoeff_valid[4], ccoeff_valid[4];
    double  zcoeff_used[4], tcoeff_used[4], ccoeff_used[4];   /* self.?coeff actually applied */
} axctd_drop_summary;

/* One profile frame (parse.py:41-92) with its calibrated values (parse.py:113-134). */
typedef struct axctd_frame {
    int64_t edge_index;         /* PCM index paired with the frame's first bit */
    uint32_t word;              /* the 32 frame bits, MSB first (binListToHex) */
    int32_t  chunk;             /* run() iteration that parsed it */
    int32_t  cint, tint;        /* parse.py:106-107 */
    int32_t  keep;              /* 1 if the row survives QC and the spike filter (:569-609) */
    int32_t  hex_returned;      /* 1 if its hex string reaches self.hexframes (:612,:323) */
    double time_s;              /* rounded, + firstpointtime        :560 */
    double depth, temperature, conductivity, salinity;   /* rounded  :561-564 */
    double r400, r7500;         /* rounded                          :565-566 */
    double time_raw, depth_raw, temperature_raw, conductivity_raw, salinity_raw, r400_raw, r7500_raw;
} axctd_frame;

/* Compact result record of one profile frame: what AXCTD_Processor exposes per frame (the lists extended at
 * AXCTDprocessor.py:315-323), with the np.round(v, 2) values (:559-566) carried as integer hundredths
 * q = rint(100 v), so that q / 100.0 on the host is bit-identical to the rounded double. */
#define AXCTD_ROW_KEEP 1            /* row survives QC and the spike filter (:569-609) */
#define AXCTD_ROW_HEX  2            /* its hex string reaches self.hexframes (:612) */
#define AXCTD_ROW_WIDE 4            /* a value does not fit the int32 hundredths: fetch the drop with axctd_batch_frames */
#define AXCTD_ROW_NAN  (-2147483647 - 1)   /* 32-bit fields */
#define AXCTD_ROW_NAN16 (-32768)          /* 16-bit fields */
typedef struct axctd_row {      /* 24 bytes */
    uint32_t word;              /* the 32 frame bits, MSB first */
    int32_t  time_c, depth_c;   /* hundredths of a second / metre */
    int16_t  temperature_c, conductivity_c, salinity_c, r400_c, r7500_c;
    uint16_t flags;             /* AXCTD_ROW_* */
} axctd_row;

/* One run() iteration (AXCTDprocessor.py:283-338). */
typedef struct axctd_chunk {
    int64_t s, e;               /* demodbufferstartind, e           :293-304 */
    int32_t status;             /* self.status after the iteration */
    int32_t n_power_total;      /* len(power_inds) after the iteration */
    int32_t n_bits;             /* bits demodulated in this iteration, -1 if none */
    int32_t first_edge, last_edge;   /* chunk-relative, demodulate.py:85,104 */
    int32_t n_head_edges;       /* edges taken from the exact zero-state recomputation */
    int32_t n_rows, n_hex;
    double  scale;              /* high_bit_scale used                :411 */
    int64_t profstartind;       /* self.profstartind after the iteration (:401, :405), -1 before the trigger */
} axctd_chunk;

typedef struct axctd_engine axctd_engine;
typedef struct axctd_batch  axctd_batch;

/* ---- engine ----------------------------------------------------------- */
int  axctd_abi_version(void);
/* 1 if the library was built with the CUDA kernels (always for the product). */
int  axctd_has_cuda(void);
/* sizeof() of the ABI structs: 0 config_desc, 1 drop_summary, 2 frame, 3 chunk, 4 row (binding self-check). */
int  axctd_struct_size(int which);
int  axctd_engine_create(int device, axctd_engine** out);
void axctd_engine_destroy(axctd_engine* e);
const char* axctd_last_error(axctd_engine* e);
/* Tunables (set before the configs / batches they are to affect).  Decomposition: "segment_len" (samples per lane of the
 * continuous pass, 0 = sized for whole waves), "seg_target" (segment length that sizing aims at, 16384), "exact_head",
 * "zc_div".  Numerics: "guard", "bit_tol", "hist_tol", "force_exact", "bitfix_all".  Kernel variants, all held to the
 * reference by the tests: "filter_variant", "ws", "fir_first", "bulk", "pair_launch" (demodulation pass); "tone_direct",
 * "tone_mma", "tone_int8", "tone_complement" (tone levels); "fuse_bits", "nosync", "max_fixups", "scan_only".
 * Scheduling of several engines on one GPU: "heavy_chain", "heavy_prio".  Block cache: "pool", "pool_trim",
 * "pool_poison".  Test hooks: "inject_misspec", "demod_probe".  Unknown names return AXCTD_ERR_ARG. */
int  axctd_engine_set_option(axctd_engine* e, const char* name, double value);
/* Run all engine work on a caller-owned CUDA stream (cudaStream_t passed as void*), so that the
 * caller can bracket it with its own events.  The engine does not take ownership. */
int  axctd_engine_set_stream(axctd_engine* e, void* cuda_stream);
/* Number of kernels launched by the engine since creation (for bench.py). */
int64_t axctd_engine_launch_count(axctd_engine* e);

/* Replaces AXCTD_Processor.initialize_AXCTD_vars + load_AXCTD_settings
 * (AXCTDprocessor.py:117-182, 212-262): copies the tables to the device. */
int  axctd_config_create(axctd_engine* e, const axctd_config_desc* desc, int* config_id);

/* ---- batch of independent drops --------------------------------------- */
/* Allocates device storage for n_drops mono int16 recordings (n_samples counts the RAW samples; a drop whose
 * config has decimate == 2 is halved on the device and reports numpoints = ceil(n/2)). */
int  axctd_batch_create(axctd_engine* e, int n_drops, const int64_t* n_samples,
                        const int32_t* config_id, axctd_batch** out);
void axctd_batch_destroy(axctd_batch* b);
/* Replaces the PCM part of readAXCTDwavfile (AXCTDprocessor.py:41-57): copies
 * one drop's int16 samples host->device (stream ordered). */
int  axctd_batch_upload(axctd_batch* b, int drop, const int16_t* pcm, int64_t n);
/* Same for a multi-channel recording as scipy.io.wavfile.read returns it (frames of `channels` interleaved int16
 * samples): the frames are copied as they are and the first channel is picked on the device
 * (AXCTDprocessor.py:46-52, `audiostream = snd[:,0]`).  n_frames counts frames. */
int  axctd_batch_upload_interleaved(axctd_batch* b, int drop, const int16_t* frames, int64_t n_frames, int channels);
/* Recordings whose samples are not 8 / 16-bit integers (24 / 32-bit PCM and float WAV files, which
 * scipy.io.wavfile.read also returns, AXCTDprocessor.py:41): the host forms (x - mean) / max|x| in double precision
 * exactly as AXCTDprocessor.py:55-57 does (and halves recordings above 50 kHz as :60-62 does) and hands the result
 * over; the drop's config must have decimate == 3 and n is the length of that signal.  The int16 statistics
 * (pcm_sum, pcm_ampl) of such a drop are not defined. */
int  axctd_batch_upload_f64(axctd_batch* b, int drop, const double* samples, int64_t n);
/* Fill a drop from samples that are already on the device: n samples of drop src_drop of batch src (same GPU), from
 * sample src_offset on.  This is how the segments of a long recording (segment.py: one upload, many drops) and the
 * points of a parameter sweep over one archive reach their batch without a second host->device copy. */
int  axctd_batch_copy_from(axctd_batch* b, int drop, axctd_batch* src, int src_drop, int64_t src_offset, int64_t n);
/* Device pointer of a drop's PCM (for callers that fill it on the GPU). */
int  axctd_batch_device_pcm(axctd_batch* b, int drop, void** dptr);
/* Replaces AXCTD_Processor.run() (AXCTDprocessor.py:267-338) for every drop
 * of the batch.  Blocks until the results are on the host. */
int  axctd_batch_run(axctd_batch* b);
/* Same, but returns after enqueueing the device work; pair with _finish. */
int  axctd_batch_run_async(axctd_batch* b);
int  axctd_batch_finish(axctd_batch* b);
/* ---- a growing recording, decoded as it arrives ------------------------------------------------------------
 * The reference's run() is a loop over 2 s iterations with a `keepgoing` flag and per-iteration result lists
 * (AXCTDprocessor.py:119, :283-338, :612), the shape a live receiver needs.  These three calls run that loop one
 * batch of new iterations at a time: the drops of `b` were created with the most samples they can take
 * (axctd_batch_create), _append adds samples as they arrive (only these cross the PCIe link), and every _run decodes
 * the iterations that have become complete -- start + pointsperloop inside the data, so the end-of-file cut of
 * :299-300 cannot apply to them -- on top of the device state the earlier runs left: tone block sums, filtering,
 * crossings and bit windows cover the new samples only, edges and bit decisions the new iterations only.  The last
 * run (final_run != 0) applies the end-of-file rules of :295-300 to what is left.  After each run the usual getters
 * (axctd_batch_summary / _rows / _chunks / _frames / _bits / _edges / _power) report everything decoded so far.
 * readAXCTDwavfile normalises with the mean and the peak of the WHOLE file (AXCTDprocessor.py:55-57), which a live
 * decoder cannot know: the caller fixes them (dc[i], ampl[i] per drop, e.g. from the first seconds or from the
 * receiver's scaling), and the result is the reference's for the recording normalised with those two numbers.
 * Recordings above 50 kHz (config decimate == 2) cannot be streamed: AXCTDprocessor.py:60-62 filters backwards. */
int  axctd_batch_stream_begin(axctd_batch* b, const double* dc, const double* ampl);
int  axctd_batch_stream_append(axctd_batch* b, int drop, const int16_t* pcm, int64_t n);
int  axctd_batch_stream_run(axctd_batch* b, int final_run);
/* Device milliseconds of the last run (CUDA events on the engine stream),
 * and of the dominant filter kernel inside it. */
int  axctd_batch_timing(axctd_batch* b, double* total_ms, double* filter_ms, double* tone_ms);

/* Device milliseconds of the five phases of the last run, in order: ingest (statistics, tone block sums, /2
 * decimation), demodulation pass, crossing bookkeeping, 400 Hz pulse search, everything after it (chunk chain, bits,
 * headers, frames, calibration). */
int  axctd_batch_phase_ms(axctd_batch* b, double* ms5);
int  axctd_batch_summary(axctd_batch* b, int drop, axctd_drop_summary* out);
/* Caller-owned buffers; each returns the number of items written or a
 * negative AXCTD_ERR_* code.  cap is the buffer capacity in items. */
/* Compact per-frame results (downloaded by axctd_batch_finish into pinned host memory). */
int64_t axctd_batch_rows(axctd_batch* b, int drop, axctd_row* out, int64_t cap);
/* Full per-frame records including the unrounded values (copied device->host on demand). */
int64_t axctd_batch_frames(axctd_batch* b, int drop, axctd_frame* out, int64_t cap);
int64_t axctd_batch_chunks(axctd_batch* b, int drop, axctd_chunk* out, int64_t cap);
int64_t axctd_batch_bits(axctd_batch* b, int drop, uint8_t* bits, double* conf, int64_t cap);
int64_t axctd_batch_edges(axctd_batch* b, int drop, int64_t* edges, double* r400, double* r7500, int64_t cap);
int64_t axctd_batch_power(axctd_batch* b, int drop, int64_t* power_inds, double* r400, double* r7500, int64_t cap);

/* ---- bench / test tooling (not part of the reference's path) --------------
 * Device-side twin of synth.py: fills one drop of the batch with a synthetic
 * AXCTD recording (BASELINE.json configs) without staging it through the host. */
typedef struct axctd_synth_desc {
    int64_t n_total, n0, tone_start, fs;
    uint64_t key1, key2;
    double nscale, gain, tone_amp;
    double sin_coef[9];
    const uint8_t* bits;        /* [nslots] transmitted bit per 1/800 s slot */
    const uint8_t* gate;        /* [nslots] carrier on/off */
    const uint8_t* parity;      /* [nslots] parity of the ones before the slot */
    int64_t nslots;
} axctd_synth_desc;
int  axctd_synth_fill(axctd_batch* b, int drop, const axctd_synth_desc* desc);
/* Copy one drop's PCM device->host (to hand device-generated audio to the CPU baseline). */
int  axctd_batch_download(axctd_batch* b, int drop, int16_t* pcm, int64_t n);
/* Evaluate the device's practical-salinity routine (gsw.SP_from_C as called at parse.py:132) and the cubic
 * conversion (parse.dataconvert, parse.py:297-301) for n caller-supplied points: sp[i] = SP_from_C(cond[i],
 * temp[i], pres[i]); poly[i] = dataconvert(cond[i], coeff) when coeff != NULL.  Known-answer tests only. */
int  axctd_calib_eval(axctd_engine* e, const double* cond, const double* temp, const double* pres, int n,
                      const double* coeff4, double* sp, double* poly);

#ifdef __cplusplus
}
#endif
#endif /* AXCTD_H */
