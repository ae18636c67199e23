"""Host-side driver of the CUDA engine: settings -> device tables, batches of
drops -> results.  Mirrors what reference AXCTD_Processor.__init__ derives from
its settings (AXCTDprocessor.py:117-262) and hands the arithmetic to the kernels
through the C ABI (include/axctd.h)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
from scipy import signal

from . import _lib

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

# reference AXCTDprocessor.py:187-208
DEFAULT_SETTINGS = {
    "minr400": 2.0, "mindr7500": 1.5, "deadfreq": 3000, "triggerrange": ([30, -1],),
    "mark_space_freqs": [400, 800], "bitrate": 800, "bit_inset": 1, "phase_error": 25,
    "usebandpass": False, "refreshrate": 2.0,
    "zcoeff_axctd": [0.72, 2.76124, -0.000238007, 0], "tcoeff_axctd": [-0.053328, 0.994372, 0.0, 0.0],
    "ccoeff_axctd": [-0.0622192, 1.04584, 0.0, 0.0], "tlims_axctd": [-10, 50], "slims_axctd": [-1, 100],
}

STATUS_EXCEPTIONS = {
    16: (IndexError, "index 0 is out of bounds for axis 0 with size 0"),          # demodulate.py:85
    17: (TypeError, "slice indices must be integers or None or have an __index__ method"),   # AXCTDprocessor.py:331 -> :304
    18: (ValueError, "operands could not be broadcast together"),                  # demodulate.py:101
    19: (ValueError, "invalid literal for int() with base 10"),                    # parse.py:278
    20: (ValueError, "zero-size array to reduction operation minimum which has no identity"),   # demodulate.py:149
    21: (IndexError, "index 0 is out of bounds for axis 0 with size 0"),           # AXCTDprocessor.py:462/:546
    32: (RuntimeError, "engine capacity exceeded"),
    33: (RuntimeError, "decision inside the numerical guard band"),
    34: (RuntimeError, "chunk chain did not converge"),
}


def load_temp_lut(path: str | None = None) -> np.ndarray:
    """parse.read_temp_LUT (reference parse.py:139-147).  A text table in the
    reference's ``idx, value`` format is used when given; otherwise the packaged
    binary copy of the same 4096 calibration values."""
    if path is not None:
        lut = []
        with open(path) as f:
            for line in f.readlines():
                c = line.strip().split(",")
                if len(c) >= 2:
                    lut.append(float(c[1]))
        return np.asarray(lut, dtype=np.float64)
    return np.fromfile(os.path.join(DATA_DIR, "temp_lut_f64le.bin"), dtype="<f8")


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class RateConfig:
    """Everything the reference derives from f_s and its settings dict."""
    fs: float
    settings: dict
    triggerrange: list
    temp_lut: np.ndarray
    desc: _lib.ConfigDesc = None
    keep: list = field(default_factory=list)
    config_id: int = -1
    decimate: int = 1            # 2: the recording is above 50 kHz and is halved on the device (fs is already f_s/2)

    def key(self):
        s = self.settings
        return (float(self.fs), float(s["minr400"]), float(s["mindr7500"]), float(s["deadfreq"]),
                tuple(float(x) for x in s["mark_space_freqs"]), bool(s["usebandpass"]), float(s["refreshrate"]),
                tuple(self.triggerrange), tuple(s["zcoeff_axctd"]), tuple(s["tcoeff_axctd"]), tuple(s["ccoeff_axctd"]),
                tuple(s["tlims_axctd"]), tuple(s["slims_axctd"]), self.temp_lut.tobytes(), int(self.decimate))

    def build(self):
        f_s, st = self.fs, self.settings
        d = _lib.ConfigDesc()
        d.fs = float(f_s)
        d.n_power = int(f_s / 10)                                        # :153
        d.d_pcm = int(np.round(f_s / 25))                                # :155
        bitrate, bit_inset, phase_error = 800, 1, 25                     # :164-166 (hard-coded)
        N = int(np.round(f_s / bitrate * (1 - phase_error / 100)))       # :170
        d.npcm = N - 2 * bit_inset                                       # :171
        d.chunk_len = int(st["refreshrate"] * f_s)                       # :222
        d.pad, d.bit_inset, d.bitrate = 100, bit_inset, bitrate          # :158
        if st["usebandpass"]:                                            # :254-257
            sos = signal.butter(6, [100, 1200], btype="bandpass", fs=f_s, output="sos")
        else:
            sos = signal.butter(6, 1200, btype="lowpass", fs=f_s, output="sos")
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        d.n_sections = sos.shape[0]
        for i in range(sos.shape[0]):
            for j in range(6):
                d.sos[i][j] = float(sos[i, j])
        d.max_pole_radius = float(max(np.abs(np.roots([1.0, r[4], r[5]])).max() for r in sos))
        f1, f2 = st["mark_space_freqs"][0], st["mark_space_freqs"][1]    # :241-242
        R = 256
        while R < d.npcm + 1:
            R *= 2
        n = np.arange(0, R + 1)
        trig1 = 2 * np.pi * n / f_s * f1                                 # :245
        trig2 = 2 * np.pi * n / f_s * f2                                 # :246
        bit_cs = np.ascontiguousarray(np.stack([np.cos(trig1), np.sin(trig1), np.cos(trig2), np.sin(trig2)], axis=1))
        npw = np.arange(0, d.n_power)
        th = [2 * np.pi * npw / f_s * f for f in (400, 7500, st["deadfreq"])]   # :260-262
        tone_cs = np.ascontiguousarray(np.stack([g(t) for t in th for g in (np.cos, np.sin)], axis=1))
        edges = np.arange(0.0, 3, 0.01)                                  # demodulate.py:130
        centers = edges[:-1] + np.diff(edges) / 2                        # demodulate.py:132
        lut = np.ascontiguousarray(self.temp_lut, dtype=np.float64)
        self.keep = [bit_cs, tone_cs, edges, centers, lut]
        d.bit_cs, d.bit_cs_len, d.tone_cs = _dptr(bit_cs), R + 1, _dptr(tone_cs)
        d.min_r400, d.min_dr7500 = float(st["minr400"]), float(st["mindr7500"])   # :225-227
        d.trigger_from_s, d.trigger_to_s = float(self.triggerrange[0]), float(self.triggerrange[1])
        d.high_bit_scale0 = 1.5                                          # :161
        for i in range(4):
            d.zcoeff[i] = float(st["zcoeff_axctd"][i]); d.tcoeff[i] = float(st["tcoeff_axctd"][i])
            d.ccoeff[i] = float(st["ccoeff_axctd"][i])
        for i in range(2):
            d.tlims[i] = float(st["tlims_axctd"][i]); d.slims[i] = float(st["slims_axctd"][i])
        d.temp_lut, d.lut_len = _dptr(lut), len(lut)
        d.hist_edges, d.hist_centers, d.n_hist_edges = _dptr(edges), _dptr(centers), len(edges)
        d.decimate = int(self.decimate)
        if self.decimate == 2:                                           # AXCTDprocessor.py:60-62 -> scipy.signal.decimate(pcm, 2)
            dsos = np.ascontiguousarray(signal.cheby1(8, 0.05, 0.8 / 2, output="sos"), dtype=np.float64)
            dzi = signal.sosfilt_zi(dsos)
            d.decim_sections = dsos.shape[0]
            ntheta = 2 * dsos.shape[0] + 1 - min((dsos[:, 2] == 0).sum(), (dsos[:, 5] == 0).sum())
            d.decim_padlen = int(3 * ntheta)                             # sosfiltfilt's default padlen
            for i in range(dsos.shape[0]):
                for j in range(6):
                    d.decim_sos[i][j] = float(dsos[i, j])
                d.decim_zi[i][0], d.decim_zi[i][1] = float(dzi[i, 0]), float(dzi[i, 1])
            d.decim_pole_radius = float(max(np.abs(np.roots([1.0, r[4], r[5]])).max() for r in dsos))
        self.desc = d
        return self


TABLE_DT = np.dtype([("word", np.uint32), ("keep", np.int32), ("hex_returned", np.int32), ("time_s", np.float64),
                     ("depth", np.float64), ("temperature", np.float64), ("conductivity", np.float64),
                     ("salinity", np.float64), ("r400", np.float64), ("r7500", np.float64)])


@dataclass
class DropResult:
    summary: _lib.DropSummary
    rows: np.ndarray            # structured array of axctd_row (compact: rounded values as integer hundredths)
    chunks: np.ndarray          # structured array of axctd_chunk
    config: RateConfig
    frames: np.ndarray = None   # structured array of axctd_frame (full records; fetched with full=True)

    def table(self) -> np.ndarray:
        """Per-frame results as the reference exposes them (rounded doubles): from the compact rows,
        q / 100.0 being bit-identical to np.round(v, 2); from the full records if a value did not fit."""
        out = np.zeros(len(self.rows), dtype=TABLE_DT)
        if len(self.rows) and (self.rows["flags"] & _lib.ROW_WIDE).any():
            if self.frames is None:
                raise RuntimeError("a value exceeds the compact row range: fetch the result with full=True")
            for k in TABLE_DT.names:
                out[k] = self.frames[k]
            return out
        out["word"] = self.rows["word"]
        out["keep"] = (self.rows["flags"] & _lib.ROW_KEEP) != 0
        out["hex_returned"] = (self.rows["flags"] & _lib.ROW_HEX) != 0
        for k in ("time_s", "depth", "temperature", "conductivity", "salinity", "r400", "r7500"):
            q = self.rows[{"time_s": "time_c"}.get(k, k + "_c")]
            v = q / 100.0
            v[q == (_lib.ROW_NAN if q.dtype == np.int32 else _lib.ROW_NAN16)] = np.nan
            out[k] = v
        return out

    @property
    def status(self):
        return int(self.summary.status)

    def raise_for_status(self):
        if self.status:
            exc, msg = STATUS_EXCEPTIONS.get(self.status, (RuntimeError, f"engine status {self.status}"))
            raise exc(msg)


def _np_dtype(struct):
    return np.dtype([(n, t) for n, t in struct._fields_], align=True)


FRAME_DT = _np_dtype(_lib.Frame)
CHUNK_DT = _np_dtype(_lib.Chunk)
ROW_DT = _np_dtype(_lib.Row)


class Engine:
    """One engine per GPU (include/axctd.h).  ``lib`` is for dependency
    injection by the test-suite's host emulation only; the default is the
    in-tree CUDA library and nothing else."""

    def __init__(self, device: int = 0, lib=None, allow_emulation: bool = False):
        self.lib = lib if lib is not None else _lib.load()
        if not self.lib.axctd_has_cuda() and not allow_emulation:
            raise RuntimeError("refusing to run without the CUDA kernels (no CPU fallback in the product)")
        h = C.c_void_p()
        rc = self.lib.axctd_engine_create(device, C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError(f"axctd_engine_create failed (rc={rc}): no usable CUDA device")
        self.h = h
        self.device = device
        self._configs = {}
        self._temp_lut = None

    def close(self):
        if getattr(self, "h", None):
            # batches that are still open hold device blocks of this engine: release them first (a Batch that is
            # garbage-collected after its engine would otherwise hand a dangling engine pointer to the library)
            for b in list(getattr(self, "_live_batches", ())):
                b.close()
            self.lib.axctd_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self) -> str:
        return (self.lib.axctd_last_error(self.h) or b"").decode()

    def set_option(self, name: str, value: float):
        if self.lib.axctd_engine_set_option(self.h, name.encode(), float(value)) != 0:
            raise ValueError(f"unknown engine option {name}")

    def set_stream(self, cuda_stream: int):
        """Run on a caller-owned CUDA stream (e.g. torch.cuda.Stream().cuda_stream)."""
        if self.lib.axctd_engine_set_stream(self.h, C.c_void_p(cuda_stream)) != 0:
            raise RuntimeError("axctd_engine_set_stream failed")

    @property
    def launch_count(self) -> int:
        return int(self.lib.axctd_engine_launch_count(self.h))

    def config(self, fs, settings=None, triggerrange=None, temp_lut=None, decimate=1) -> RateConfig:
        """Rate class for recordings of effective rate ``fs``.  decimate=2: the batch receives the raw
        recording sampled at 2*fs and halves it on the device first."""
        st = {k: (list(v) if isinstance(v, list) else v) for k, v in DEFAULT_SETTINGS.items()}
        for k, v in (settings or {}).items():
            st[k] = v
        if temp_lut is None:
            if self._temp_lut is None:
                self._temp_lut = load_temp_lut()
            temp_lut = self._temp_lut
        rc = RateConfig(fs=fs, settings=st, triggerrange=list(triggerrange) if triggerrange is not None else [30, -1],
                        temp_lut=np.asarray(temp_lut, dtype=np.float64), decimate=int(decimate))
        key = rc.key()
        if key in self._configs:
            return self._configs[key]
        rc.build()
        cid = C.c_int32(-1)
        r = self.lib.axctd_config_create(self.h, C.byref(rc.desc), C.byref(cid))
        if r != 0:
            raise RuntimeError(f"axctd_config_create failed ({r}): {self.error()}")
        rc.config_id = cid.value
        self._configs[key] = rc
        return rc

    def calib_eval(self, cond, temp, pres, coeff=None):
        """Known-answer hook: the device's SP_from_C (parse.py:132) and, with ``coeff``, dataconvert(cond, coeff)
        (parse.py:297-301) for arrays of points.  Returns (sp, poly or None)."""
        c = np.ascontiguousarray(cond, dtype=np.float64).reshape(-1)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(temp, dtype=np.float64), c.shape))
        p = np.ascontiguousarray(np.broadcast_to(np.asarray(pres, dtype=np.float64), c.shape))
        sp = np.empty_like(c)
        poly = np.empty_like(c) if coeff is not None else None
        cf = np.ascontiguousarray(coeff, dtype=np.float64) if coeff is not None else None
        rc = self.lib.axctd_calib_eval(self.h, c.ctypes.data, t.ctypes.data, p.ctypes.data, len(c),
                                       cf.ctypes.data if cf is not None else None, sp.ctypes.data,
                                       poly.ctypes.data if poly is not None else None)
        if rc != 0:
            raise RuntimeError(f"axctd_calib_eval failed ({rc}): {self.error()}")
        return sp, poly

    def batch(self, n_samples, configs) -> "Batch":
        return Batch(self, n_samples, configs)

    def process(self, pcm_list, configs) -> list:
        """Upload, run and collect a batch of mono int16 recordings."""
        b = self.batch([len(p) for p in pcm_list], configs)
        try:
            for i, p in enumerate(pcm_list):
                b.upload(i, p)
            b.run()
            return [b.result(i) for i in range(len(pcm_list))]
        finally:
            b.close()


class Batch:
    def __init__(self, eng: Engine, n_samples, configs):
        self.eng = eng
        self.lib = eng.lib
        self.n = len(n_samples)
        self.configs = list(configs)
        ns = (C.c_int64 * self.n)(*[int(x) for x in n_samples])
        ci = (C.c_int32 * self.n)(*[c.config_id for c in self.configs])
        h = C.c_void_p()
        rc = self.lib.axctd_batch_create(eng.h, self.n, ns, ci, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"axctd_batch_create failed ({rc}): {eng.error()}")
        self.h = h
        self.n_samples = [int(x) for x in n_samples]
        if not hasattr(eng, "_live_batches"):
            import weakref
            eng._live_batches = weakref.WeakSet()
        eng._live_batches.add(self)

    def close(self):
        if getattr(self, "h", None):
            self.lib.axctd_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.eng.error()}")

    def upload(self, i: int, pcm: np.ndarray):
        """Mono samples (n,), or frames (n, channels) as scipy.io.wavfile.read returns them: the first channel
        is picked on the device (AXCTDprocessor.py:46-52)."""
        if isinstance(pcm, np.ndarray) and pcm.dtype == np.float64 and self.configs[i].decimate == 3:
            # the normalised signal of a recording with wide samples (24 / 32-bit, float: AXCTDprocessor.normalised_signal)
            pcm = np.ascontiguousarray(pcm)
            self._check(self.lib.axctd_batch_upload_f64(self.h, i, pcm.ctypes.data, pcm.size), "axctd_batch_upload_f64")
            return
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if pcm.ndim == 2:
            self._check(self.lib.axctd_batch_upload_interleaved(self.h, i, pcm.ctypes.data, pcm.shape[0], pcm.shape[1]),
                        "axctd_batch_upload_interleaved")
            return
        self._check(self.lib.axctd_batch_upload(self.h, i, pcm.ctypes.data, pcm.size), "axctd_batch_upload")

    def upload_ptr(self, i: int, host_ptr: int, n: int):
        self._check(self.lib.axctd_batch_upload(self.h, i, host_ptr, n), "axctd_batch_upload")

    def copy_from(self, i: int, src: "Batch", src_drop: int, src_offset: int, n: int):
        """Fill drop i with n samples of ``src``'s drop ``src_drop`` (already on the device) from ``src_offset`` on."""
        self._check(self.lib.axctd_batch_copy_from(self.h, i, src.h, src_drop, int(src_offset), int(n)), "axctd_batch_copy_from")

    # ---- a growing recording (include/axctd.h, "decoded as it arrives")
    def stream_begin(self, dc, ampl):
        """Switch the batch to streaming: every drop starts empty (its length at creation is the most it can take)
        and is normalised with the given (dc, ampl) instead of whole-file statistics."""
        dc = np.ascontiguousarray(np.broadcast_to(np.asarray(dc, dtype=np.float64), (self.n,)))
        am = np.ascontiguousarray(np.broadcast_to(np.asarray(ampl, dtype=np.float64), (self.n,)))
        self._check(self.lib.axctd_batch_stream_begin(self.h, _dptr(dc), _dptr(am)), "axctd_batch_stream_begin")

    def stream_append(self, i: int, pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        self._check(self.lib.axctd_batch_stream_append(self.h, i, pcm.ctypes.data, pcm.size), "axctd_batch_stream_append")

    def stream_run(self, final: bool = False):
        self._check(self.lib.axctd_batch_stream_run(self.h, 1 if final else 0), "axctd_batch_stream_run")

    def device_ptr(self, i: int) -> int:
        p = C.c_void_p()
        self._check(self.lib.axctd_batch_device_pcm(self.h, i, C.byref(p)), "axctd_batch_device_pcm")
        return p.value

    def run(self):
        self._check(self.lib.axctd_batch_run(self.h), "axctd_batch_run")

    def run_async(self):
        self._check(self.lib.axctd_batch_run_async(self.h), "axctd_batch_run_async")

    def finish(self):
        self._check(self.lib.axctd_batch_finish(self.h), "axctd_batch_finish")

    def timing(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.axctd_batch_timing(self.h, C.byref(a), C.byref(b), C.byref(c)), "axctd_batch_timing")
        ph = (C.c_double * 5)()
        self._check(self.lib.axctd_batch_phase_ms(self.h, ph), "axctd_batch_phase_ms")
        return dict(total_ms=a.value, filter_ms=b.value, tone_ms=c.value, ingest_ms=ph[0], demod_ms=ph[1],
                    crossings_ms=ph[2], search_ms=ph[3], decode_ms=ph[4])

    def summary(self, i: int) -> _lib.DropSummary:
        s = _lib.DropSummary()
        self._check(self.lib.axctd_batch_summary(self.h, i, C.byref(s)), "axctd_batch_summary")
        return s

    def result(self, i: int, full: bool = True) -> DropResult:
        """Results of drop i.  full=False skips the device->host copy of the full per-frame records
        (unrounded values): the compact rows already hold everything the reference's API exposes."""
        s = self.summary(i)
        nrow = int(self.lib.axctd_batch_rows(self.h, i, None, 0))
        rows = np.zeros(max(nrow, 0), dtype=ROW_DT)
        if len(rows):
            n = self.lib.axctd_batch_rows(self.h, i, rows.ctypes.data, len(rows))
            assert n == len(rows), n
        chunks = np.zeros(max(int(s.n_chunks), 0), dtype=CHUNK_DT)
        if len(chunks):
            n = self.lib.axctd_batch_chunks(self.h, i, chunks.ctypes.data, len(chunks))
            chunks = chunks[:max(n, 0)]
        frames = None
        if full:
            frames = np.zeros(max(int(s.n_frames), 0), dtype=FRAME_DT)
            if len(frames):
                n = self.lib.axctd_batch_frames(self.h, i, frames.ctypes.data, len(frames))
                assert n == len(frames), n
        return DropResult(summary=s, rows=rows, chunks=chunks, config=self.configs[i], frames=frames)

    def synth_fill(self, i: int, spec):
        """bench / test tooling: generate synth.DropSpec ``spec`` directly in device memory."""
        import synth
        n_total, truth = synth.build_bitplan(spec)
        assert n_total == self.n_samples[i], (n_total, self.n_samples[i])
        par = np.zeros(len(truth.bits) + 1, dtype=np.int64)
        np.cumsum(truth.bits, out=par[1:])
        par = np.ascontiguousarray((par[:-1] & 1).astype(np.uint8))
        bits = np.ascontiguousarray(truth.bits, dtype=np.uint8)
        gate = np.ascontiguousarray(truth.gate, dtype=np.uint8)
        d = _lib.SynthDesc()
        d.n_total, d.n0, d.tone_start, d.fs = n_total, truth.n0, truth.tone_start_sample, spec.fs
        d.key1, d.key2 = synth.stream_key(spec.seed, 1), synth.stream_key(spec.seed, 2)
        d.nscale, d.gain, d.tone_amp = synth.noise_sigma(spec) / synth._IH8_SIGMA, synth.gain(spec), spec.tone_amp
        for k in range(9):
            d.sin_coef[k] = synth._SIN_COEF[k]
        d.bits, d.gate, d.parity, d.nslots = bits.ctypes.data, gate.ctypes.data, par.ctypes.data, len(bits)
        self._check(self.lib.axctd_synth_fill(self.h, i, C.byref(d)), "axctd_synth_fill")
        return truth

    def download(self, i: int) -> np.ndarray:
        out = np.empty(self.n_samples[i], dtype=np.int16)
        self._check(self.lib.axctd_batch_download(self.h, i, out.ctypes.data, out.size), "axctd_batch_download")
        return out

    def bits(self, i: int):
        n = int(self.summary(i).n_bits)
        bits = np.zeros(n, dtype=np.uint8)
        conf = np.zeros(n, dtype=np.float64)
        if n:
            r = self.lib.axctd_batch_bits(self.h, i, bits.ctypes.data, conf.ctypes.data, n)
            assert r == n, r
        return bits, conf

    def edges(self, i: int):
        n = int(self.summary(i).n_edges)
        e = np.zeros(n, dtype=np.int64)
        a = np.zeros(n, dtype=np.float64)
        b = np.zeros(n, dtype=np.float64)
        if n:
            r = self.lib.axctd_batch_edges(self.h, i, e.ctypes.data, a.ctypes.data, b.ctypes.data, n)
            assert r == n, r
        return e, a, b

    def power(self, i: int):
        n = int(self.summary(i).n_power)
        p = np.zeros(n, dtype=np.int64)
        a = np.zeros(n, dtype=np.float64)
        b = np.zeros(n, dtype=np.float64)
        if n:
            r = self.lib.axctd_batch_power(self.h, i, p.ctypes.data, a.ctypes.data, b.ctypes.data, n)
            assert r == n, r
        return p, a, b
