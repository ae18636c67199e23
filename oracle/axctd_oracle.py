"""TEST INFRASTRUCTURE (oracle) -- numpy restatement of the AXCTD hot path.

This module is the CHECKER for the CUDA engine.  It restates, chunk for chunk,
what the reference's pure-Python path computes (reference files are cited as
``file:line`` into the upstream tree) using numpy / scipy on the CPU.  It is
imported only by tests/, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
cpu_baseline / ``--impl reference`` legs; the product package never imports it
and has no CPU fallback.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4.1).  The restatement is pinned instead against (a) the one
known-answer frame of reference README.md:87 and the constants of SURVEY.md
section 4.2 (tests/test_oracle_units.py) and (b) outputs of the UNMODIFIED
reference run in the build container through oracle/ref_shim.py, committed as
fixtures under tests/golden/ by oracle/make_golden.py (tests/test_oracle_golden.py).
Salinity (gsw.SP_from_C, gsw==3.3.1 in the reference's requirements.txt) is not available
here; it is restated from the published PSS-78 algorithm in oracle/pss78.py and pinned to the
UNESCO check value and to the six check values of the GSW documentation for gsw_SP_from_C
(tests/test_oracle_units.py); the Hill extension below SP = 2 has no published vector and
stays PARITY UNPINNED.

Differences from the reference that do not change results: the per-bit and
per-window single-bin DFTs, the nearest-power-sample lookup and the CRC scan are
vectorised (summation order differs at the 1e-16 level from numpy's pairwise
``np.sum`` of complex128); control flow, chunk chain, index bookkeeping
(including the n+1 edge quirk, reference AXCTDprocessor.py:413-429) and every
threshold are kept as shipped.
"""
from __future__ import annotations

import os

import numpy as np
from scipy import signal
from scipy.io import wavfile

try:  # as a package (tests) or as a plain directory on sys.path (tools)
    from . import pss78
except ImportError:  # pragma: no cover
    import pss78

_HERE = os.path.dirname(os.path.abspath(__file__))
_LUT_BIN = os.path.join(os.path.dirname(_HERE), "axctdprocessor_b200", "data", "temp_lut_f64le.bin")


def load_temp_lut(path=None):
    """reference parse.py:139-147 (text table) -- the committed copy is the same
    4096 float64 values in binary form."""
    if path is not None and os.path.isfile(path) and path.endswith(".txt"):
        lut = []
        with open(path) as f:
            for line in f:
                c = line.strip().split(",")
                if len(c) >= 2:
                    lut.append(float(c[1]))
        return lut
    return np.fromfile(_LUT_BIN, dtype="<f8").tolist()


# --------------------------------------------------------------------------
# ingest -- reference AXCTDprocessor.py:38-73
# --------------------------------------------------------------------------
def normalise_pcm(snd, fs):
    snd = np.asarray(snd)
    if snd.ndim == 1:
        a = snd
    elif snd.ndim == 2:
        a = snd[:, 0]                                   # :50 first channel only
    else:
        raise Exception("Too many dimensions for an audio file!")
    dc = np.mean(a)                                     # :55
    ampl = np.max(np.abs(a))                            # :56 (int16 abs wraps at -32768)
    with np.errstate(all="ignore"):
        pcm = (a.astype(float) - dc) / ampl             # :57
    if fs > 50000:                                      # :60-62
        pcm = signal.decimate(pcm, 2)
        fs /= 2
    return pcm, fs


def read_wav(path):
    fs, snd = wavfile.read(path)                        # :41
    return normalise_pcm(snd, fs)


# --------------------------------------------------------------------------
# protocol helpers -- reference parse.py
# --------------------------------------------------------------------------
_CRC_DIV = (1, 1, 0, 0, 1, 0, 1)                         # parse.py:312


def check_crc(bits):
    """parse.py:310-322: remainder of the 32-bit word modulo 1100101 is zero."""
    r = [int(bool(b)) for b in bits]
    for k in range(26):
        if r[k]:
            for i in range(7):
                r[k + i] ^= _CRC_DIV[i]
    return not sum(r)


def bits_to_int(bits):                                  # parse.py:331-340
    x = 0
    for b in bits:
        x = (x << 1) | (1 if b else 0)
    return x


def bits_to_hex(bits):                                  # parse.py:363-379
    return "".join("0123456789abcdef"[bits_to_int(bits[i:i + 4])] for i in range(0, len(bits) - 3, 4))


def crc_valid_positions(bits):
    """Vectorised ``frame[0:2]==[1,0] and check_crc(frame)`` for every start
    position with 32 bits available (parse.py:68, :224)."""
    b = np.asarray(bits, dtype=np.uint8)
    n = len(b) - 31
    if n <= 0:
        return np.zeros(0, dtype=bool)
    win = np.lib.stride_tricks.sliding_window_view(b, 32)[:n].astype(np.uint8).copy()
    div = np.array(_CRC_DIV, dtype=np.uint8)
    for k in range(26):
        m = win[:, k].astype(bool)
        win[m, k:k + 7] ^= div
    ok = win.sum(axis=1) == 0
    return ok & (b[:n] == 1) & (b[1:n + 1] == 0)


def dataconvert(x, coeff):                              # parse.py:297-301
    out = 0
    for i, c in enumerate(coeff):
        out = out + c * x ** i
    return out


def trim_header(bits_in):
    """parse.py:157-183."""
    bits = list(bits_in)
    bits[:25] = [True] * 25
    last_index_pulse = 0
    ones25 = 0
    for i, b in enumerate(bits):
        if b:
            ones25 += 1
            if i > 10 and sum(1 for v in bits[i - 7:i + 1] if v) == 8:
                last_index_pulse = i
        if i > 24:
            if bits[i - 25]:
                ones25 -= 1
            if i >= 400 and ones25 <= 20:
                break
    return bits[last_index_pulse:last_index_pulse + 32 * 75]


def init_metadata():                                    # parse.py:187-192
    return {"tcoeff": [0, 1, 0, 0], "ccoeff": [0, 1, 0, 0], "zcoeff": [1, 1, 1, 1],
            "serial_no": None, "probe_code": None, "max_depth": None, "misc": None,
            "tcoeff_hex": ["", "", "", ""], "ccoeff_hex": ["", "", "", ""], "zcoeff_hex": ["", "", "", ""],
            "tcoeff_valid": [False] * 4, "ccoeff_valid": [False] * 4, "zcoeff_valid": [False] * 4}


def coefficient_from_hex(text):
    """parse.py:277-278 (python float arithmetic; ValueError on A/C/E/F)."""
    chex = text.upper().replace("B", "+").replace("D", "-")
    return int(chex[:9]) / 1e7 * 10 ** int(chex[9:])


def parse_header(bits):
    """parse.py:197-285."""
    bits = [1 if b else 0 for b in bits]
    counter_found = [False] * 72
    md = init_metadata()
    frame_data = [None] * 72
    lastframe = -1
    n = len(bits)
    s = 0
    while lastframe < 71 and s < n - 32:
        if bits[s:s + 2] != [1, 0] or not check_crc(bits[s:s + 32]):
            s += 1
        else:
            cb = bits[s + 2:s + 10]
            cur = bits_to_int(cb[5:]) + 64 if cb[:5] == [1, 1, 1, 1, 1] else bits_to_int(cb)
            if cur <= 71:
                counter_found[cur] = True
                lastframe = cur
                frame_data[cur] = bits_to_hex(bits[s + 10:s + 26])
            s += 32
    if sum(counter_found[4:6]) == 2:
        md["serial_no"] = frame_data[4] + frame_data[5]
    if counter_found[6]:
        md["max_depth"] = frame_data[6]
    if counter_found[7]:
        md["probe_code"] = frame_data[7]
    for key, top in (("zcoeff_hex", 21), ("tcoeff_hex", 33), ("ccoeff_hex", 45)):
        for i, cf in enumerate(range(top, top - 10, -3)):
            if sum(counter_found[cf:cf + 3]) == 3:
                md[key][i] = "".join(frame_data[cf:cf + 3])
    for c in ("t", "c", "z"):
        for i in range(4):
            if md[c + "coeff_hex"][i] != "":
                md[c + "coeff"][i] = coefficient_from_hex(md[c + "coeff_hex"][i])
                md[c + "coeff_valid"][i] = True
    md["frame_data"] = frame_data
    md["counter_found"] = counter_found
    return md


# --------------------------------------------------------------------------
# DSP helpers -- reference demodulate.py
# --------------------------------------------------------------------------
def boxsmooth_lag(data, window, startind):
    """demodulate.py:39-48.  Reads the INPUT array, so entries before startind
    (already smoothed by earlier calls) are mixed with raw new entries."""
    data = np.asarray(data, dtype=np.float64)
    out = data.copy()
    n = len(data)
    for i in range(startind, min(n, window)):
        out[i] = _nanmean_seq(data[0:i + 1])
    lo = max(startind, window)
    if n > lo:
        cols = [data[lo - window + j: n - window + j] for j in range(window + 1)]
        nan = [np.isnan(c) for c in cols]
        acc = np.where(nan[0], 0.0, cols[0])
        cnt = (~nan[0]).astype(np.float64)
        for j in range(1, window + 1):
            acc = acc + np.where(nan[j], 0.0, cols[j])
            cnt = cnt + (~nan[j])
        with np.errstate(all="ignore"):
            out[lo:] = acc / cnt
    return out


def _nanmean_seq(v):
    tot, cnt = 0.0, 0
    for x in v:
        if not np.isnan(x):
            tot += float(x)
            cnt += 1
    return tot / cnt if cnt else np.nan


def demodulate_axctd(pcm, fs, edge_buffer, sos, bitrate, trig1, trig2, Npcm, bit_inset, high_bit_scale):
    """demodulate.py:59-116."""
    pcmlow = signal.sosfilt(sos, pcm)                   # :74
    sgn = pcmlow >= 0                                   # :77-78 (sign, zero -> +1); NaN -> False
    if np.isnan(pcmlow).any():
        raise FloatingPointError("NaN in filtered PCM")
    zc = np.flatnonzero(sgn[:-1] != sgn[1:])            # :79
    zc = zc[zc >= edge_buffer]                          # :82
    if len(zc) == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")   # :85
    z = zc.tolist()
    nz = len(z)
    edges = [z[0]]
    step = fs / bitrate
    c = 0
    while c < nz - 5:                                   # :90-93
        target = z[c] + step
        best, bj = None, 0
        for j in range(4):
            d = abs(z[c + 1 + j] - target)
            if best is None or d < best:
                best, bj = d, j
        c += 1 + bj
        edges.append(z[c])
    e = np.asarray(edges[:-1], dtype=np.int64)
    if len(e) and e[-1] + bit_inset + Npcm > len(pcmlow):
        raise ValueError("operands could not be broadcast together")          # :101 short window
    idx = e[:, None] + bit_inset + np.arange(Npcm)[None, :]
    w = pcmlow[idx] if len(e) else np.zeros((0, Npcm))
    s1 = np.abs((w * np.cos(trig1)).sum(axis=1) + 1j * (w * np.sin(trig1)).sum(axis=1))       # :101
    s2 = np.abs((w * np.cos(trig2)).sum(axis=1) + 1j * (w * np.sin(trig2)).sum(axis=1)) * high_bit_scale  # :102
    next_ind = edges[-1] - 1                            # :104
    with np.errstate(all="ignore"):
        conf = s2 / s1                                  # :110
    bits = (s1 >= s2).astype(np.uint8)                  # :111-114
    return bits, conf, np.asarray(edges, dtype=np.int64), next_ind


def adjust_scale_factor(confs, scale_factor):
    """demodulate.py:124-157."""
    npts = len(confs)
    confs = np.asarray(confs)
    bin_edges = np.arange(0.0, 3, 0.01)
    dist, bin_edges = np.histogram(confs, bins=bin_edges)
    centers = bin_edges[:-1] + np.diff(bin_edges) / 2
    cumpct = 100 * np.cumsum(dist) / npts
    slope = np.array((cumpct[1] - cumpct[0]) / (centers[1] - centers[0]))
    slope = np.append(slope, (cumpct[2:] - cumpct[:-2]) / (centers[2:] - centers[:-2]))
    slope = np.append(slope, (cumpct[-1] - cumpct[-2]) / (centers[-1] - centers[-2]))
    in_range = (cumpct >= 30) & (cumpct <= 65)
    centers = centers[in_range]
    slope = slope[in_range]
    mn = np.min(slope)
    ismin = np.where(slope == mn)[0]
    thr = np.nanmean([centers[ismin[0]], centers[ismin[-1]]])
    return scale_factor / thr


# --------------------------------------------------------------------------
# processor -- reference AXCTDprocessor.py:80-627
# --------------------------------------------------------------------------
DEFAULT_SETTINGS = {                                    # :187-208
    "minr400": 2.0, "mindr7500": 1.5, "deadfreq": 3000, "mark_space_freqs": [400, 800],
    "bitrate": 800, "bit_inset": 1, "phase_error": 25, "usebandpass": False, "refreshrate": 2.0,
    "zcoeff_axctd": [0.72, 2.76124, -0.000238007, 0], "tcoeff_axctd": [-0.053328, 0.994372, 0.0, 0.0],
    "ccoeff_axctd": [-0.0622192, 1.04584, 0.0, 0.0], "tlims_axctd": [-10, 50], "slims_axctd": [-1, 100],
}


class OracleProcessor:
    """Restatement of reference class AXCTD_Processor.  ``settings`` uses the
    processor's internal key names (AXCTDprocessor.py:191-208); ``{}`` is the
    faithful as-shipped behaviour.  ``triggerrange`` overrides the hard-coded
    [30, -1] (AXCTDprocessor.py:250-251) for wired mode."""

    def __init__(self, pcm, fs, settings=None, triggerrange=None, temp_lut=None, keep_trace=True):
        self.audiostream = np.asarray(pcm, dtype=np.float64)
        self.f_s = fs
        self.numpoints = len(self.audiostream)
        self.settings = {k: (list(v) if isinstance(v, list) else v) for k, v in DEFAULT_SETTINGS.items()}
        for k, v in (settings or {}).items():
            self.settings[k] = v
        st = self.settings
        self.tempLUT = list(temp_lut) if temp_lut is not None else load_temp_lut()
        self.metadata = init_metadata()
        self.metadata["counter_found_2"] = [False] * 72
        self.metadata["counter_found_3"] = [False] * 72
        f_s = self.f_s
        # :133-171
        self.p400 = np.array([]); self.p7500 = np.array([]); self.pdead = np.array([])
        self.r400 = np.array([]); self.r7500 = np.array([])
        self.power_inds = []
        self.firstpulse400 = -1
        self.profstartind = -1
        self.firstpointtime = -1
        self.mean7500pwr = np.nan
        self.N_power = int(f_s / 10)
        self.power_smooth_window = 5
        self.d_pcm = int(np.round(f_s / 25))
        self.demod_Npad = 100
        self.next_demod_ind = 0
        self.high_bit_scale = 1.5
        self.bitrate = 800
        self.bit_inset = 1
        self.phase_error = 25
        N = int(np.round(f_s / self.bitrate * (1 - self.phase_error / 100)))
        self.Npcm = N - 2 * self.bit_inset
        self.binary_buffer = []
        self.binary_buffer_inds = []
        self.binary_buffer_conf = []
        self.r400_buffer = []
        self.r7500_buffer = []
        self.past_headers = False
        self.header1_read = self.header2_read = self.header3_read = False
        # :222-262
        self.minpointsperloop = int(st["refreshrate"] * f_s)
        self.minR400 = st["minr400"]; self.minR400_inprof = st["minr400"] / 2
        self.mindR7500 = st["mindr7500"]; self.mindR7500_inprof = self.mindR7500 / 2
        self.deadfreq = st["deadfreq"]
        self.zcoeff = st["zcoeff_axctd"]; self.tcoeff = st["tcoeff_axctd"]; self.ccoeff = st["ccoeff_axctd"]
        self.tlims = st["tlims_axctd"]; self.slims = st["slims_axctd"]
        self.f1, self.f2 = st["mark_space_freqs"][0], st["mark_space_freqs"][1]
        self.trig1 = 2 * np.pi * np.arange(0, self.Npcm) / f_s * self.f1
        self.trig2 = 2 * np.pi * np.arange(0, self.Npcm) / f_s * self.f2
        self.triggerrange = list(triggerrange) if triggerrange is not None else [30, -1]
        if st["usebandpass"]:
            self.sos_filter = signal.butter(6, [100, 1200], btype="bandpass", fs=f_s, output="sos")
        else:
            self.sos_filter = signal.butter(6, 1200, btype="lowpass", fs=f_s, output="sos")
        self.theta400 = 2 * np.pi * np.arange(0, self.N_power) / f_s * 400
        self.theta7500 = 2 * np.pi * np.arange(0, self.N_power) / f_s * 7500
        self.thetadead = 2 * np.pi * np.arange(0, self.N_power) / f_s * self.deadfreq
        self._E = np.stack([np.cos(t) + 1j * np.sin(t) for t in (self.theta400, self.theta7500, self.thetadead)], axis=1)
        self.time = []; self.r400_prof = []; self.r7500_prof = []; self.hexframes = []
        self.depth = []; self.temperature = []; self.conductivity = []; self.salinity = []
        self.status = -1
        self.keep_trace = keep_trace
        self.trace = []
        self.all_bits = []; self.all_edges = []; self.all_conf = []
        self.all_frames = []       # (global edge index, hex, Cint, Tint) of every CRC-valid profile frame
        self.all_raw = []          # (time, z, T, C, S, r400, r7500) of the same frames BEFORE rounding (parse.py:92)
        self.chunk_starts = []

    # -- main loop, AXCTDprocessor.py:267-338
    def run(self):
        self.status = 0
        self.demodbufferstartind = 0
        while True:
            e = self.demodbufferstartind + self.minpointsperloop        # :293
            if self.numpoints - self.demodbufferstartind < 4 * self.N_power:   # :295
                break
            elif e >= self.numpoints:                                    # :299-300
                e = self.numpoints - 1
            if not isinstance(self.demodbufferstartind, (int, np.integer)):
                raise TypeError("slice indices must be integers")       # :331 float start index
            self.demod_buffer = self.audiostream[self.demodbufferstartind:e]   # :304
            data = self.iterate(e)
            if len(data) > 1:                                            # :315-323
                self.time.extend(data[1]); self.r400_prof.extend(data[2]); self.r7500_prof.extend(data[3])
                self.depth.extend(data[4]); self.temperature.extend(data[5]); self.conductivity.extend(data[6])
                self.salinity.extend(data[7]); self.hexframes.extend(data[8])
            if self.status > 0:                                          # :327-331
                if self.next_demod_ind > self.demod_Npad:
                    self.demodbufferstartind += self.next_demod_ind - self.demod_Npad
                else:
                    self.demodbufferstartind += self.f_s / self.bitrate
            else:
                self.demodbufferstartind = e                             # :333
        return self

    # -- one chunk, AXCTDprocessor.py:346-627
    def iterate(self, e):
        s0 = self.demodbufferstartind
        f_s = self.f_s
        pstart = len(self.power_inds)
        new_inds = list(range(s0, e - self.N_power, self.d_pcm))         # :357
        self.power_inds.extend(new_inds)
        if new_inds:                                                     # :358-364
            rel = np.asarray(new_inds) - s0
            win = np.lib.stride_tricks.sliding_window_view(self.demod_buffer, self.N_power)[rel]
            P = np.abs(win @ self._E)
            self.p400 = np.append(self.p400, P[:, 0]); self.p7500 = np.append(self.p7500, P[:, 1])
            self.pdead = np.append(self.pdead, P[:, 2])
        self.p400 = boxsmooth_lag(self.p400, self.power_smooth_window, pstart)     # :367-369
        self.p7500 = boxsmooth_lag(self.p7500, self.power_smooth_window, pstart)
        self.pdead = boxsmooth_lag(self.pdead, self.power_smooth_window, pstart)
        with np.errstate(all="ignore"):
            self.r400 = np.append(self.r400, np.log10(self.p400[pstart:] / self.pdead[pstart:]))    # :370
            self.r7500 = np.append(self.r7500, np.log10(self.p7500[pstart:] / self.pdead[pstart:]))  # :371
        if self.status == 0:                                             # :375-380
            m = np.where(self.r400[pstart:] >= self.minR400)[0]
            if len(m) > 0:
                self.firstpulse400 = self.power_inds[pstart:][m[0]]
                self.status = 1
        rec = dict(s=int(s0), e=int(e), status=self.status)
        if self.status >= 1:
            if self.power_inds[-1] >= self.firstpulse400 + int(f_s * 5.5) and np.isnan(self.mean7500pwr):   # :388-393
                pia = np.asarray(self.power_inds)
                s75 = np.argmin(np.abs(self.firstpulse400 + int(f_s * 4.5) - pia))
                e75 = np.argmin(np.abs(self.firstpulse400 + int(f_s * 5.5) - pia))
                with np.errstate(all="ignore"):
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        self.mean7500pwr = np.nanmean(self.r7500[s75:e75])
            if self.power_inds[-1] > self.firstpulse400 + int(self.triggerrange[0] * f_s):     # :397-408
                if not np.isnan(self.mean7500pwr) and self.status == 1:
                    m = np.where(self.r7500[pstart:] - self.mean7500pwr >= self.mindR7500)[0]
                    if len(m) > 0:
                        self.profstartind = self.power_inds[pstart:][m[0]]
                        self.status = 2
                elif self.triggerrange[1] > 0 and self.power_inds[-1] >= self.firstpulse400 + int(f_s * self.triggerrange[1]):
                    self.profstartind = self.firstpulse400 + int(f_s * self.triggerrange[1])
                    self.status = 2
                if self.profstartind > 0 and self.firstpointtime <= 0:
                    self.firstpointtime = self.profstartind / f_s
            bits, conf, edges, self.next_demod_ind = demodulate_axctd(                       # :411
                self.demod_buffer, f_s, self.demod_Npad, self.sos_filter, self.bitrate, self.trig1, self.trig2,
                self.Npcm, self.bit_inset, self.high_bit_scale)
            self.binary_buffer.extend(bits.tolist())                                           # :413
            new_bit_inds = edges + s0                                                          # :415
            self.binary_buffer_inds.extend(new_bit_inds.tolist())
            self.binary_buffer_conf.extend(conf.tolist())
            rp = np.asarray(self.power_inds[pstart:])
            near = np.argmin(np.abs(rp[None, :] - new_bit_inds[:, None]), axis=1)             # :425,428
            self.r400_buffer.extend(self.r400[pstart:][near].tolist())
            self.r7500_buffer.extend((self.r7500[pstart:][near] - self.mean7500pwr).tolist())
            rec.update(nbits=len(bits), first_edge=int(edges[0]), last_edge=int(edges[-1]),
                       next_ind=int(self.next_demod_ind), scale=float(self.high_bit_scale))
            if self.keep_trace:
                self.all_bits.extend(bits.tolist()); self.all_edges.extend(new_bit_inds.tolist())
                self.all_conf.extend(conf.tolist())
        if self.status >= 1 and not self.past_headers:
            self._headers()
        data = [self.status]
        if self.status == 2:
            data = self._profile()
        rec.update(n_power=len(self.power_inds), nrows=(len(data[1]) if len(data) > 1 else 0),
                   nhex=(len(data[8]) if len(data) > 1 else 0), status=self.status, profstart=int(self.profstartind))
        self.trace.append(rec)
        return data

    # -- header handling, AXCTDprocessor.py:433-535
    def _headers(self):
        f_s = self.f_s
        headerdata = [None, None]
        firstbin = self.binary_buffer_inds[0]
        lastbin = self.binary_buffer_inds[-1]
        arr = np.asarray(self.binary_buffer_inds)
        fp = self.firstpulse400
        p1s, p1e = fp + int(f_s * 2.3), fp + int(f_s * 3.3)
        p2s, p2e = fp + int(f_s * 10.5), fp + int(f_s * 14.8)
        p3s, p3e = fp + int(f_s * 20), fp + int(f_s * 24.5)
        half = int(f_s * 0.5)
        if firstbin <= p1s and lastbin >= p1e and not self.header1_read:            # :459-468
            a = np.where(arr >= p1s - half)[0][0]
            b = np.where(arr <= p1e + half)[0][-1]
            self.high_bit_scale = adjust_scale_factor(self.binary_buffer_conf[a:b], self.high_bit_scale)
            self.header1_read = True
        for slot, (ps_, pe_, flag) in enumerate(((p2s, p2e, "header2_read"), (p3s, p3e, "header3_read"))):
            if firstbin <= ps_ and lastbin >= pe_ and not getattr(self, flag):       # :472-501
                a = np.where(arr >= ps_ - half)[0][0]
                b = np.where(arr <= pe_ + half)[0][-1]
                hb = trim_header(self.binary_buffer[a:b])
                if len(hb) >= 72 * 32:
                    headerdata[slot] = parse_header(hb)
                    setattr(self, flag, True)
        for i, header in enumerate(headerdata):                                       # :506-524
            if header is not None:
                self.metadata[f"frame_data_{i + 2}"] = header["frame_data"]
                self.metadata[f"counter_found_{i + 2}"] = header["counter_found"]
                for c in ("t", "c", "z"):
                    for ci in range(4):
                        if header[c + "coeff_valid"][ci]:
                            self.metadata[c + "coeff"][ci] = header[c + "coeff"][ci]
                            self.metadata[c + "coeff_hex"][ci] = header[c + "coeff_hex"][ci]
                            self.metadata[c + "coeff_valid"][ci] = True
                for key in ("serial_no", "probe_code", "max_depth", "misc"):
                    if header[key] is not None and self.metadata[key] is None:
                        self.metadata[key] = header[key]
        if headerdata[0] is not None or headerdata[1] is not None:                    # :529-535
            if sum(self.metadata["tcoeff_valid"]) == 4:
                self.tcoeff = self.metadata["tcoeff"]
            if sum(self.metadata["ccoeff_valid"]) == 4:
                self.ccoeff = self.metadata["ccoeff"]
            if sum(self.metadata["tcoeff_valid"]) == 4:      # (sic) z is guarded by the T flag
                self.zcoeff = self.metadata["zcoeff"]

    # -- profile parsing, AXCTDprocessor.py:540-627 and parse.py:41-134
    def _profile(self):
        f_s = self.f_s
        self.past_headers = True
        if self.binary_buffer_inds[0] <= self.profstartind:                            # :545-551
            first = int(np.where(np.asarray(self.binary_buffer_inds) > self.profstartind)[0][0])
            self.binary_buffer = self.binary_buffer[first:]
            self.binary_buffer_inds = self.binary_buffer_inds[first:]
            self.binary_buffer_conf = self.binary_buffer_conf[first:]
            self.r400_buffer = self.r400_buffer[first:]
            self.r7500_buffer = self.r7500_buffer[first:]
        inds = np.asarray(self.binary_buffer_inds)
        binbufftimes = (inds - self.profstartind) / f_s                                # :554
        bits = self.binary_buffer
        numbits = len(bits)
        valid = crc_valid_positions(bits)
        r75 = np.asarray(self.r7500_buffer, dtype=np.float64)
        r40 = np.asarray(self.r400_buffer, dtype=np.float64)
        nv = len(valid)
        with np.errstate(all="ignore"):
            valid = valid & (r75[:nv] > 0)                                             # parse.py:68
        cand = np.flatnonzero(valid)
        hexframes, pos = [], []
        s = 0
        while s < numbits - 32:                                                        # parse.py:57-89
            j = np.searchsorted(cand, s)
            if j >= len(cand) or cand[j] >= numbits - 32:
                s = max(s, numbits - 32)
                break
            s = int(cand[j])
            pos.append(s)
            s += 32
        next_buffer_ind = s
        pos = np.asarray(pos, dtype=np.int64)
        barr = np.asarray(bits, dtype=np.int64)
        if len(pos):
            w = barr[pos[:, None] + np.arange(32)[None, :]]
            pw = 1 << np.arange(11, -1, -1)
            cint = (w[:, 2:14] * pw).sum(axis=1)                                       # parse.py:107
            tint = (w[:, 14:26] * pw).sum(axis=1)                                      # parse.py:106
            nib = (w.reshape(len(pos), 8, 4) * np.array([8, 4, 2, 1])).sum(axis=2)
            hexframes = ["".join("0123456789abcdef"[v] for v in row) for row in nib]
            times = binbufftimes[pos]
            lut = np.asarray(self.tempLUT, dtype=np.float64)
            tun = np.where((tint >= 0) & (tint <= len(lut) - 1), lut[np.clip(tint, 0, len(lut) - 1)], np.nan)
            cun = cint * 60 / 4096                                                     # parse.py:125
            z = dataconvert(times, self.zcoeff)                                        # parse.py:117
            T = dataconvert(tun, self.tcoeff)
            C = dataconvert(cun, self.ccoeff)
            S = pss78.SP_from_C(C, T, z)                                               # parse.py:132
            r400 = r40[pos]; r7500 = r75[pos]
            if self.keep_trace:
                for k in range(len(pos)):
                    self.all_frames.append((int(inds[pos[k]]), hexframes[k], int(cint[k]), int(tint[k])))
                    self.all_raw.append((float(times[k]), float(z[k]), float(T[k]), float(C[k]), float(S[k]),
                                         float(r400[k]), float(r7500[k])))
        else:
            times = z = T = C = S = r400 = r7500 = np.zeros(0)
        # :560-566
        times = np.round(np.asarray(times) + self.firstpointtime, 2)
        depths = np.round(z, 2); temps = np.round(T, 2); conds = np.round(C, 2); psals = np.round(S, 2)
        r400 = np.round(r400, 2); r7500 = np.round(r7500, 2)
        with np.errstate(all="ignore"):
            bad = ((r7500 < self.mindR7500_inprof) | (r400 < self.minR400_inprof) | (temps < self.tlims[0])
                   | (temps > self.tlims[1]) | (psals < self.slims[0]) | (psals > self.slims[1]))     # :572-574
        good = ~bad
        times, depths, temps, conds, psals, r400, r7500 = (a[good] for a in (times, depths, temps, conds, psals, r400, r7500))
        data = [self.status]
        if len(temps) > 0:                                                              # :587-613
            thresh, off = 10, 35
            with np.errstate(all="ignore"):
                Tm = np.percentile(temps, 50)
                Tlo = Tm - thresh * (Tm - np.percentile(temps, 50 - off))
                Thi = Tm + thresh * (np.percentile(temps, 50 + off) - Tm)
                Sm = np.percentile(psals, 50)
                Slo = Sm - thresh * (Sm - np.percentile(psals, 50 - off))
                Shi = Sm + thresh * (np.percentile(psals, 50 + off) - Sm)
                bad = (temps < Tlo) | (temps > Thi) | (psals < Slo) | (psals > Shi)
            good = ~bad
            times, depths, temps, conds, psals, r400, r7500 = (a[good] for a in (times, depths, temps, conds, psals, r400, r7500))
            if len(temps) > 0:
                data = [self.status, times, r400, r7500, depths, temps, conds, psals, hexframes]
        # :618-621
        self.binary_buffer = self.binary_buffer[next_buffer_ind:]
        self.binary_buffer_inds = self.binary_buffer_inds[next_buffer_ind:]
        self.r400_buffer = self.r400_buffer[next_buffer_ind:]
        self.r7500_buffer = self.r7500_buffer[next_buffer_ind:]
        return data


# --------------------------------------------------------------------------
# output file -- reference processAXCTD.py:143-183
# --------------------------------------------------------------------------
def format_output(ap, wavfile_name, timerange, settings):
    """Text of the output file for a finished processor (KeyError 'zcoeff_default'
    when any coefficient set is incomplete, as shipped: processAXCTD.py:161-167)."""
    out = []
    fs = ap.f_s
    out.append(f"AXCTD profile for {wavfile_name}\n")
    out.append(f"Sampling frequency (fs): {fs} Hz\n")
    out.append(f"Audio file length: {ap.numpoints / fs} sec\n")
    out.append(f"400 Hz pulse start: {ap.firstpulse400 / fs} sec\n")
    out.append(f"7500 Hz tone start: {ap.profstartind / fs} sec\n")
    out.append("\nAXCTD header information:\n")
    for desc, key in zip(["Probe Code", "Maximum Depth (m)", "Probe Serial"], ["probe_code", "max_depth", "serial_no"]):
        out.append(f"{desc}: {ap.metadata[key]}\n")
    out.append("Conversion equations:\n")
    for coeff, desc, symb in zip(["z", "t", "c"], ["Depth", "Temperature", "Conductivity"], ["t", "T", "C"]):
        if sum(ap.metadata[coeff + "coeff_valid"]) == 4:
            cfield, dflt = coeff + "coeff", ""
        else:
            cfield, dflt = coeff + "coeff_default", "(default)"
        eqn = " + ".join([f"{val}*{symb}^{i}" for i, val in enumerate(ap.metadata[cfield])])
        out.append(f"{desc}: {eqn} {dflt}\n")
    out.append("\nProcessor Settings:\n")
    out.append(f"Time Range: {timerange[0]} sec to {timerange[1] if timerange[1] >= 0 else 'N/A'} sec\n")
    out.append(f"Min. 400 Hz power ratio: {settings['minR400']}\n")
    out.append(f"Min. 7500 Hz power ratio: {settings['mindR7500']}\n")
    out.append(f"Dead frequency: {settings['deadfreq']}\n")
    out.append(f"Points per loop: {settings['pointsperloop']}\n")
    tr = settings["triggerrange"]
    out.append(f"Trigger range: {tr[0]} sec to {tr[1] if tr[1] >= 0 else 'N/A'} sec\n")
    out.append("\nAXCTD Profile:\n")
    out.append("Time (s), Hex Frame, Depth (m), Temperature (C), Conductivity (mS/cm), Salinity (PSU)\n")
    for (t, hf, z, T, C, S) in zip(ap.time, ap.hexframes, ap.depth, ap.temperature, ap.conductivity, ap.salinity):
        out.append(f"{t:8.2f},  {hf},{z:10.2f},{T:16.2f},{C:21.2f},{S:15.2f}\n")
    return "".join(out)


def process_pcm(snd, fs, settings=None, triggerrange=None):
    pcm, fs2 = normalise_pcm(snd, fs)
    return OracleProcessor(pcm, fs2, settings=settings, triggerrange=triggerrange).run()
