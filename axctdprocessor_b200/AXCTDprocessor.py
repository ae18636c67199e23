"""Drop-in mirror of the reference module ``AXCTDprocessor`` (reference
AXCTDprocessor.py:38-627): same function / class names, constructor arguments,
result attributes and error behaviour, with the whole demodulate -> decode path
executed by the CUDA engine (include/axctd.h) instead of numpy loops.

``mode="faithful"`` (default) reproduces the reference exactly as shipped,
including its inert settings keys and its crashes;
``mode="wired"`` makes the documented flags act (time range trimming, CLI key
names, trigger range).
"""
from __future__ import annotations

import os
import struct

import numpy as np

from . import engine as _engine

_default_engines = {}


def default_engine(device: int = 0) -> _engine.Engine:
    if device not in _default_engines:
        _default_engines[device] = _engine.Engine(device)
    return _default_engines[device]


def read_wav(path):
    """RIFF/WAVE reader returning what scipy.io.wavfile.read returns (reference AXCTDprocessor.py:41): (fs, array
    of shape (n,) or (n, channels)) with dtype uint8 -> widened to int16 here (exact), int16, int32 for 24-bit
    (left-justified, as scipy stores it) and 32-bit PCM, int64 for 64-bit PCM, float32 / float64 for IEEE files."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[0:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"File format {data[0:4]!r} not understood. Only 'RIFF' and 'RIFX' supported.")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, nch, fs, _, _, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == 0xFFFE and len(body) >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, nch, fs, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError("incomplete WAV file (missing fmt or data chunk)")
    tag, nch, fs, bits = fmt
    if nch < 1:
        raise ValueError("malformed WAV file: the fmt chunk declares no channels")
    bps = bits // 8
    nval = (len(pcm) // (bps * nch)) * nch if bps else 0
    if tag == 1 and bits == 8:
        # scipy.io.wavfile.read returns uint8 for 8-bit PCM (the reference then subtracts the mean, :55-57): every
        # value fits the engine's int16 input exactly
        a = np.frombuffer(pcm, dtype=np.uint8, count=nval).astype(np.int16)
    elif tag == 1 and bits == 16:
        a = np.frombuffer(pcm, dtype="<i2", count=nval)
    elif tag == 1 and bits == 24:
        raw = np.frombuffer(pcm, dtype=np.uint8, count=3 * nval).reshape(-1, 3)
        wide = np.zeros((nval, 4), dtype=np.uint8)
        wide[:, 1:] = raw                               # left-justified in 32 bits, as scipy stores 24-bit samples
        a = wide.view("<i4").reshape(-1)
    elif tag == 1 and bits in (32, 64):
        a = np.frombuffer(pcm, dtype="<i4" if bits == 32 else "<i8", count=nval)
    elif tag == 3 and bits in (32, 64):
        a = np.frombuffer(pcm, dtype="<f4" if bits == 32 else "<f8", count=nval)
    else:
        raise ValueError(f"Unknown wave file format: format tag {tag}, {bits} bit")
    if nch > 1:
        a = a.reshape(-1, nch)
    return fs, a


read_wav_pcm16 = read_wav          # (name of earlier rounds)


def normalised_signal(audiostream, fs):
    """AXCTDprocessor.py:55-62 on the host, for recordings whose samples are not 8 / 16-bit integers: the same numpy
    expressions as the reference ((x - mean) / max|x| in double precision; scipy.signal.decimate(pcm, 2) above
    50 kHz).  The engine takes the result as it is (axctd_batch_upload_f64) -- int16 recordings never come here,
    their normalisation and halving run on the GPU."""
    from scipy import signal
    pcm_dc = np.mean(audiostream)
    pcm_ampl = np.max(np.abs(audiostream))
    pcm = (audiostream.astype(np.float64) - pcm_dc) / pcm_ampl
    if fs > 50000:
        pcm = signal.decimate(pcm, 2)
        fs /= 2
    return np.ascontiguousarray(pcm, dtype=np.float64), fs


class WidePCM(np.ndarray):
    """Normalised double-precision signal of a recording with wide samples (what the reference's audiostream holds)."""
    decimate = 3


def _first_channel(snd):
    shape = np.shape(snd)
    if len(shape) == 1:                                  # AXCTDprocessor.py:46-52
        return np.ascontiguousarray(snd)
    if len(shape) == 2:
        return np.ascontiguousarray(snd[:, 0])
    raise Exception("Too many dimensions for an audio file!")


def _frames(snd):
    """The recording as the engine takes it: mono samples (n,) or interleaved frames (n, channels) -- the first
    channel of a multi-channel file (AXCTDprocessor.py:50) is picked on the GPU (axctd_batch_upload_interleaved)."""
    if len(np.shape(snd)) not in (1, 2):
        raise Exception("Too many dimensions for an audio file!")
    return np.ascontiguousarray(snd)


class RawPCM(np.ndarray):
    """int16 samples as read from the WAV file.  ``decimate`` = 2 marks a recording above 50 kHz: the
    engine halves it on the GPU (reference AXCTDprocessor.py:60-62), and ``len()`` / f_s of the
    processor refer to the halved signal as in the reference."""
    decimate = 1


def readAXCTDwavfile(inputfile, timerange):
    """Reference AXCTDprocessor.py:38-73.  Returns the RAW first-channel int16
    samples and f_s: normalisation ((x-mean)/max|x|, :55-57) and the /2 decimation of
    recordings above 50 kHz (:60-62, f_s becomes the float f_s/2) happen on the GPU.
    As shipped, any positive time bound raises NameError (:65-70)."""
    return _read_recording(inputfile, timerange, _first_channel)


def _read_recording(inputfile, timerange, pick):
    fs, snd = read_wav(inputfile)
    if snd.dtype != np.int16:                            # 24 / 32-bit, float: normalised (and halved) on the host
        pcm, fs = normalised_signal(_first_channel(snd), fs)
        if timerange[1] > 0 or timerange[0] > 0:
            raise NameError("name 'self' is not defined")    # :66 / :69
        return pcm.view(WidePCM), fs
    audiostream = pick(snd).view(RawPCM)
    if fs > 50000:                                       # :60-62
        audiostream.decimate = 2
        fs /= 2
    if timerange[1] > 0 or timerange[0] > 0:
        raise NameError("name 'self' is not defined")    # :66 / :69
    return audiostream, fs


def header_metadata(frame_data, counter_found):
    """The metadata half of parse.parse_header (reference parse.py:247-285)
    from the 72 decoded 16-bit payloads."""
    md = {"tcoeff": [0, 1, 0, 0], "ccoeff": [0, 1, 0, 0], "zcoeff": [1, 1, 1, 1], "serial_no": None,
          "probe_code": None, "max_depth": None, "misc": None, "tcoeff_hex": ["", "", "", ""],
          "ccoeff_hex": ["", "", "", ""], "zcoeff_hex": ["", "", "", ""], "tcoeff_valid": [False] * 4,
          "ccoeff_valid": [False] * 4, "zcoeff_valid": [False] * 4}
    found = [bool(x) for x in counter_found]
    fd = [("%04x" % int(v)) if found[i] else None for i, v in enumerate(frame_data)]
    if sum(found[4:6]) == 2:
        md["serial_no"] = fd[4] + fd[5]
    if found[6]:
        md["max_depth"] = fd[6]
    if found[7]:
        md["probe_code"] = fd[7]
    for key, top in (("zcoeff_hex", 21), ("tcoeff_hex", 33), ("ccoeff_hex", 45)):
        for i, cf in enumerate(range(top, top - 10, -3)):
            if sum(found[cf:cf + 3]) == 3:
                md[key][i] = "".join(fd[cf:cf + 3])
    for coeff in ("t", "c", "z"):
        for i in range(4):
            if md[coeff + "coeff_hex"][i] != "":
                chex = md[coeff + "coeff_hex"][i].upper().replace("B", "+").replace("D", "-")
                md[coeff + "coeff"][i] = int(chex[:9]) / 1E7 * 10 ** int(chex[9:])     # parse.py:278
                md[coeff + "coeff_valid"][i] = True
    md["frame_data"] = fd
    md["counter_found"] = found
    return md


def initialize_axctd_metadata():
    md = header_metadata([0] * 72, [False] * 72)
    del md["frame_data"], md["counter_found"]
    return md


class AXCTD_Processor:
    """Reference class AXCTD_Processor (AXCTDprocessor.py:80-627)."""

    def __init__(self, audiofile, timerange=[0, -1], user_settings={}, mode="faithful", engine=None, device=0):
        self.audiofile = audiofile
        self.mode = mode
        self._engine = engine
        self._device = device
        if mode == "wired":
            fs, snd = read_wav(audiofile)
            a = _frames(snd)
            if timerange[1] > 0:
                a = a[:int(fs * timerange[1])]
            if timerange[0] > 0:
                a = a[int(fs * timerange[0]):]
            if snd.dtype != np.int16:
                a, fs = normalised_signal(_first_channel(a), fs)
                a = a.view(WidePCM)
            else:
                a = np.ascontiguousarray(a).view(RawPCM)
                if fs > 50000:
                    a.decimate = 2
                    fs /= 2
            self.audiostream, self.f_s = a, fs
        else:
            self.audiostream, self.f_s = _read_recording(audiofile, timerange, _frames)     # (frames as read: channel 0 is picked on the GPU)
        self._decimate = int(getattr(self.audiostream, "decimate", 1))
        self.numpoints = (len(self.audiostream) + 1) // 2 if self._decimate == 2 else len(self.audiostream)
        self.init_default_AXCTD_settings()
        for csetting in user_settings:
            self.settings[csetting] = user_settings[csetting]          # :95-96 (verbatim overlay)
        self.initialize_AXCTD_vars()
        self.load_AXCTD_settings()
        self.time = []
        self.r400_prof = []
        self.r7500_prof = []
        self.hexframes = []
        self.depth = []
        self.temperature = []
        self.conductivity = []
        self.salinity = []

    # -- AXCTDprocessor.py:117-182
    def initialize_AXCTD_vars(self):
        self.keepgoing = True
        self.past_headers = False
        self.header1_read = self.header2_read = self.header3_read = False
        self.metadata = initialize_axctd_metadata()
        self.metadata["counter_found_2"] = [False] * 72
        self.metadata["counter_found_3"] = [False] * 72
        lut_path = "temp_LUT.txt" if os.path.isfile("temp_LUT.txt") else None     # :130 (cwd-relative)
        self.tempLUT = _engine.load_temp_lut(lut_path).tolist()
        self.firstpulse400 = -1
        self.profstartind = -1
        self.firstpointtime = -1
        self.firstpulsetime = -1
        self.mean7500pwr = np.nan
        self.f_s_power = 25
        self.N_power = int(self.f_s / 10)
        self.power_smooth_window = 5
        self.d_pcm = int(np.round(self.f_s / self.f_s_power))
        self.demod_Npad = 100
        self.high_bit_scale = 1.5
        self.bitrate, self.bit_inset, self.phase_error = 800, 1, 25
        N = int(np.round(self.f_s / self.bitrate * (1 - self.phase_error / 100)))
        self.Npcm = N - 2 * self.bit_inset
        self.status = -1

    # -- AXCTDprocessor.py:187-208
    def init_default_AXCTD_settings(self):
        self.settings = {k: (list(v) if isinstance(v, list) else v) for k, v in _engine.DEFAULT_SETTINGS.items()}

    # -- AXCTDprocessor.py:212-262
    def load_AXCTD_settings(self):
        st = self.settings
        self.minpointsperloop = int(st["refreshrate"] * self.f_s)
        self.minR400 = st["minr400"]
        self.minR400_inprof = st["minr400"] / 2
        self.mindR7500 = st["mindr7500"]
        self.mindR7500_inprof = self.mindR7500 / 2
        self.deadfreq = st["deadfreq"]
        self.zcoeff = st["zcoeff_axctd"]
        self.tcoeff = st["tcoeff_axctd"]
        self.ccoeff = st["ccoeff_axctd"]
        self.tlims = st["tlims_axctd"]
        self.slims = st["slims_axctd"]
        self.f1 = st["mark_space_freqs"][0]
        self.f2 = st["mark_space_freqs"][1]
        self.triggerrange = [30, -1]                       # :250 (hard-coded; set the attribute to override)

    def _engine_settings(self):
        st = dict(self.settings)
        st.update(minr400=self.minR400, mindr7500=self.mindR7500, deadfreq=self.deadfreq,
                  zcoeff_axctd=list(self.zcoeff), tcoeff_axctd=list(self.tcoeff), ccoeff_axctd=list(self.ccoeff),
                  tlims_axctd=list(self.tlims), slims_axctd=list(self.slims), mark_space_freqs=[self.f1, self.f2])
        return st

    # -- AXCTDprocessor.py:267-338
    def run(self):
        self.maxtime = self.numpoints / self.f_s
        self.status = 0
        eng = self._engine or default_engine(self._device)
        cfg = eng.config(self.f_s, settings=self._engine_settings(), triggerrange=self.triggerrange,
                         temp_lut=np.asarray(self.tempLUT, dtype=np.float64), decimate=self._decimate)
        b = eng.batch([len(self.audiostream)], [cfg])
        try:
            b.upload(0, self.audiostream)
            b.run()
            res = b.result(0)
            self._collect(res)
            self._timing = b.timing()
        finally:
            b.close()
        print("[+] Processing status: 100%")
        self.keepgoing = False
        if self.mode != "lenient":
            res.raise_for_status()
        return self

    def _collect(self, res):
        s = res.summary
        self.result = res
        self.firstpulse400 = int(s.firstpulse400)
        self.profstartind = int(s.profstartind)
        self.firstpulsetime = self.firstpulse400 / self.f_s if self.firstpulse400 >= 0 else -1
        self.firstpointtime = float(s.firstpointtime) if s.profstartind > 0 else -1
        self.mean7500pwr = float(s.mean7500pwr)
        self.high_bit_scale = float(s.high_bit_scale)
        self.status = int(res.chunks["status"][-1]) if len(res.chunks) else 0
        self.header1_read, self.header2_read, self.header3_read = (bool(x) for x in s.header_read)
        self.past_headers = s.profile_chunk >= 0
        # header metadata merge, AXCTDprocessor.py:505-535
        for slot in range(2):
            if not s.header_parsed[slot]:
                continue
            try:
                header = header_metadata(list(s.frame_data[slot]), list(s.counter_found[slot]))
            except ValueError:
                if self.mode == "lenient":
                    continue
                raise
            self.metadata[f"frame_data_{slot + 2}"] = header["frame_data"]
            self.metadata[f"counter_found_{slot + 2}"] = header["counter_found"]
            for coeff in ("t", "c", "z"):
                for ci in range(4):
                    if header[coeff + "coeff_valid"][ci]:
                        self.metadata[coeff + "coeff"][ci] = header[coeff + "coeff"][ci]
                        self.metadata[coeff + "coeff_hex"][ci] = header[coeff + "coeff_hex"][ci]
                        self.metadata[coeff + "coeff_valid"][ci] = True
            for key in ("serial_no", "probe_code", "max_depth", "misc"):
                if header[key] is not None and self.metadata[key] is None:
                    self.metadata[key] = header[key]
        if s.header_parsed[0] or s.header_parsed[1]:
            if sum(self.metadata["tcoeff_valid"]) == 4:
                self.tcoeff = self.metadata["tcoeff"]
            if sum(self.metadata["ccoeff_valid"]) == 4:
                self.ccoeff = self.metadata["ccoeff"]
            if sum(self.metadata["tcoeff_valid"]) == 4:     # (sic) AXCTDprocessor.py:534
                self.zcoeff = self.metadata["zcoeff"]
        fr = res.table()
        kept = fr[fr["keep"] == 1]
        self.time = list(kept["time_s"])
        self.r400_prof = list(kept["r400"])
        self.r7500_prof = list(kept["r7500"])
        self.depth = list(kept["depth"])
        self.temperature = list(kept["temperature"])
        self.conductivity = list(kept["conductivity"])
        self.salinity = list(kept["salinity"])
        self.hexframes = ["%08x" % int(w) for w in fr["word"][fr["hex_returned"] == 1]]     # :612 (not QC filtered)
