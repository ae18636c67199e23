"""Deterministic synthetic AXCTD drop generator (bench / test tooling).

Produces mono int16 PCM that follows the signal timeline the reference
documents (reference README.md:81-96; reference AXCTDprocessor.py:436-456):
lead-in noise, 3 x [400 Hz pulse, 72-frame header, gap], back-to-back 32-bit
data frames at 800 baud (mark 400 Hz = 1, space 800 Hz = 0, phase-continuous),
a 7500 Hz profile tone, additive noise everywhere.

Every operation is integer arithmetic or a single IEEE-754 double multiply /
add / divide (no libm, no library RNG), so the PCM is bit-identical on every
platform and can be regenerated on the GPU box from a seed; tests/golden keeps
a sha256 of the PCM next to every expected output.  The CUDA generator in
axctdprocessor_b200/csrc (bench only) mirrors the same arithmetic.
"""
from __future__ import annotations

import hashlib
import math
import struct
from dataclasses import dataclass, field

import numpy as np

BITRATE = 800
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)

# sin(pi/2 * y), y in [0, 1]: odd Taylor polynomial to y^17 (error < 5e-14)
_SIN_COEF = [(-1.0) ** k * (math.pi / 2.0) ** (2 * k + 1) / math.factorial(2 * k + 1)
             for k in range(9)]


def mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    x = (x ^ (x >> np.uint64(30))) * _M1
    x = (x ^ (x >> np.uint64(27))) * _M2
    return x ^ (x >> np.uint64(31))


def stream_key(seed: int, stream: int) -> int:
    return (seed * 0x632BE59BD9B4E019 + stream * 0xD1342543DE82EF95 + 0x1234567) & (2 ** 64 - 1)


def hash_u64(seed: int, stream: int, idx: np.ndarray) -> np.ndarray:
    key = np.uint64(stream_key(seed, stream))
    with np.errstate(over="ignore"):
        return mix64(mix64(idx.astype(np.uint64) * _GOLD + key) + key)


def sin_turns(num: np.ndarray, den: int) -> np.ndarray:
    """sin(2*pi*num/den) for integer arrays 0 <= num < den, libm-free."""
    x4 = 4 * num.astype(np.int64)
    neg = x4 > 2 * den
    x4 = np.where(neg, 4 * den - x4, x4)
    x4 = np.where(x4 > den, 2 * den - x4, x4)
    y = x4.astype(np.float64) / float(den)
    y2 = y * y
    acc = np.full_like(y, _SIN_COEF[8])
    for k in range(7, -1, -1):
        acc = acc * y2 + _SIN_COEF[k]
    s = acc * y
    return np.where(neg, -s, s)


def crc6(bits26) -> list:
    """6-bit CRC, generator 1100101 (reference parse.py:312, README.md:96):
    remainder of the 26 frame bits followed by six zeros."""
    div = (1, 1, 0, 0, 1, 0, 1)
    r = list(bits26) + [0] * 6
    for k in range(26):
        if r[k]:
            for i in range(7):
                r[k + i] ^= div[i]
    return r[26:]


def _int_bits(v: int, n: int) -> list:
    return [(v >> (n - 1 - i)) & 1 for i in range(n)]


def encode_coefficient(val: float) -> str:
    """12 hex chars 'S dddddddd S dd' with B='+' and D='-' such that the
    reference's int(chars[:9])/1e7 * 10**int(chars[9:]) (parse.py:277-278)
    reproduces val to 8 significant digits."""
    if val == 0:
        return "b00000000b00"
    sgn = "b" if val > 0 else "d"
    a = abs(val)
    ex = int(math.floor(math.log10(a)))
    mant = int(round(a / 10.0 ** ex * 1e7))
    if mant >= 100000000:
        mant //= 10
        ex += 1
    esgn = "b" if ex >= 0 else "d"
    return f"{sgn}{mant:08d}{esgn}{abs(ex):02d}"


@dataclass
class DropSpec:
    fs: int = 44100
    duration_s: float = 60.0
    seed: int = 1
    snr_db: float = 40.0
    lead_in_s: float = 5.0
    pulse_s: float = 1.8
    gap_s: float = 5.0
    tone_after_pulse_s: float = 35.0
    tone_amp: float = 0.3
    full_scale: float = 0.5
    zcoeff: tuple = (0.72, 2.76124, -0.000238007, 0.0)
    tcoeff: tuple = (-0.053328, 0.994372, 0.0, 0.0)
    ccoeff: tuple = (-0.0622192, 1.04584, 0.0, 0.0)
    serial: int = 0x00123456
    max_depth_hex: int = 0x1000
    probe_code: int = 0xA000
    spike_every: int = 997       # every n-th data frame carries a temperature spike
    channels: int = 1
    mark_hz: int = 400           # tone of a 1 bit / of a 0 bit (integers); the defaults keep the PCM of every
    space_hz: int = 800          #   existing fixture bit-identical, other pairs take the general phase formula


@dataclass
class DropTruth:
    bits: np.ndarray             # transmitted bit per bit slot (uint8)
    gate: np.ndarray             # 1 where the FSK carrier is on
    n0: int                      # sample index of bit slot 0
    first_pulse_sample: int
    tone_start_sample: int
    header_bit_start: list = field(default_factory=list)
    data_bit_start: int = 0
    data_frames: np.ndarray = None   # (n,2) Cint,Tint


def header_frames(spec: DropSpec) -> list:
    """72 header frames (reference parse.py:199-203, 231-270)."""
    data = [None] * 72
    filler = hash_u64(spec.seed, 7, np.arange(72))
    for k in range(72):
        data[k] = int(filler[k] & np.uint64(0xFFFF))
    data[4] = (spec.serial >> 16) & 0xFFFF
    data[5] = spec.serial & 0xFFFF
    data[6] = spec.max_depth_hex
    data[7] = spec.probe_code
    for base, coeffs in ((12, spec.zcoeff), (24, spec.tcoeff), (36, spec.ccoeff)):
        # coefficient index i lives in frames base+9-3i .. +2 (parse.py:258-270)
        for i, c in enumerate(coeffs):
            hx = encode_coefficient(c)
            cf = base + 9 - 3 * i
            for j in range(3):
                data[cf + j] = int(hx[4 * j:4 * j + 4], 16)
    frames = []
    for k in range(72):
        cnt = _int_bits(k, 8) if k < 64 else [1, 1, 1, 1, 1] + _int_bits(k - 64, 3)
        b26 = [1, 0] + cnt + _int_bits(data[k], 16)
        frames.append(b26 + crc6(b26))
    return frames


def _profile_ints(spec: DropSpec, nframes: int):
    k = np.arange(nframes, dtype=np.int64)
    h = hash_u64(spec.seed, 11, k)
    jt = (h & np.uint64(7)).astype(np.int64) - 3
    jc = ((h >> np.uint64(8)) & np.uint64(7)).astype(np.int64) - 3
    # slow monotone drift: integer arithmetic only
    tint = 3000 - (1800 * k) // max(nframes, 18000) - (k % 400) // 40 + jt
    cint = 3700 - (1300 * k) // max(nframes, 18000) + jc
    if spec.spike_every > 0:
        spike = (k % spec.spike_every) == (spec.spike_every // 2)
        tint = np.where(spike, tint + 900, tint)
    tint = np.clip(tint, 1, 4093)
    cint = np.clip(cint, 0, 4095)
    return cint, tint


def build_bitplan(spec: DropSpec):
    fs = spec.fs
    n_total = int(round(spec.duration_s * fs))
    n0 = int(round(spec.lead_in_s * fs))
    nslots = max(int((n_total - n0) * BITRATE // fs), 0) + 2
    bits = np.ones(nslots, dtype=np.uint8)
    gate = np.zeros(nslots, dtype=np.uint8)
    pulse_bits = int(round(spec.pulse_s * BITRATE))
    gap_bits = int(round(spec.gap_s * BITRATE))
    hdr = np.array(sum(header_frames(spec), []), dtype=np.uint8)
    pos = 0
    hstarts = []
    for rep in range(3):
        if pos + pulse_bits + len(hdr) > nslots:
            break
        gate[pos:pos + pulse_bits] = 1
        pos += pulse_bits
        hstarts.append(pos)
        bits[pos:pos + len(hdr)] = hdr
        gate[pos:pos + len(hdr)] = 1
        pos += len(hdr)
        if rep < 2:
            pos += gap_bits
    data_start = pos
    nframes = max(0, (nslots - pos) // 32)
    cint, tint = _profile_ints(spec, nframes)
    if nframes:
        fb = np.zeros((nframes, 32), dtype=np.uint8)
        fb[:, 0] = 1
        for i in range(12):
            fb[:, 2 + i] = (cint >> (11 - i)) & 1
            fb[:, 14 + i] = (tint >> (11 - i)) & 1
        # CRC by the linearity of the remainder: xor of per-bit remainders
        rem = np.zeros((nframes, 6), dtype=np.uint8)
        for j in range(26):
            e = [0] * 26
            e[j] = 1
            rj = np.array(crc6(e), dtype=np.uint8)
            rem ^= (fb[:, j:j + 1] & rj[None, :])
        fb[:, 26:] = rem
        bits[pos:pos + nframes * 32] = fb.reshape(-1)
        gate[pos:pos + nframes * 32] = 1
    truth = DropTruth(bits=bits, gate=gate, n0=n0, first_pulse_sample=n0,
                      tone_start_sample=n0 + int(round(spec.tone_after_pulse_s * fs)),
                      header_bit_start=hstarts, data_bit_start=data_start,
                      data_frames=np.stack([cint, tint], axis=1) if nframes else np.zeros((0, 2), np.int64))
    return n_total, truth


def noise_sigma(spec: DropSpec) -> float:
    return math.sqrt(0.5 / (10.0 ** (spec.snr_db / 10.0)))


_IH8_SIGMA = math.sqrt(8.0 * (65536.0 ** 2 - 1.0) / 12.0)


def gain(spec: DropSpec) -> float:
    return spec.full_scale * 32767.0 / (1.0 + spec.tone_amp)


def generate_drop(spec: DropSpec, return_truth: bool = False, block: int = 1 << 20):
    """int16 PCM (n,) (or (n, channels) with the signal in channel 0)."""
    fs = spec.fs
    n_total, truth = build_bitplan(spec)
    par = np.zeros(len(truth.bits) + 1, dtype=np.int64)
    np.cumsum(truth.bits, out=par[1:])
    ones = par.copy()
    par &= 1
    nscale = noise_sigma(spec) / _IH8_SIGMA
    g = gain(spec)
    out = np.empty(n_total, dtype=np.int16)
    n0 = truth.n0
    for s in range(0, n_total, block):
        e = min(n_total, s + block)
        n = np.arange(s, e, dtype=np.int64)
        # noise: Irwin-Hall(8) of 16-bit fields from two 64-bit hashes
        h1 = hash_u64(spec.seed, 1, n)
        h2 = hash_u64(spec.seed, 2, n)
        acc = np.zeros(e - s, dtype=np.int64)
        for h in (h1, h2):
            for sh in (0, 16, 32, 48):
                acc += ((h >> np.uint64(sh)) & np.uint64(0xFFFF)).astype(np.int64)
        x = (acc.astype(np.float64) - 262140.0) * nscale
        # FSK
        rel = n - n0
        on = rel >= 0
        numer = np.where(on, rel, 0) * BITRATE
        b = numer // fs
        rem = numer - b * fs
        bit = truth.bits[b].astype(np.int64)
        if (spec.mark_hz, spec.space_hz) == (400, 800):
            q = par[b] * fs + np.where(bit == 1, 1, 2) * rem
            q = q % (2 * fs)
            fsk = sin_turns(q, 2 * fs)
        else:
            # phase in turns = [(ones before the slot * mark + zeros before it * space) * fs + f * rem] / (BITRATE * fs)
            done = (ones[b] * spec.mark_hz + (b - ones[b]) * spec.space_hz) % BITRATE
            q = (done * fs + np.where(bit == 1, spec.mark_hz, spec.space_hz) * rem) % (BITRATE * fs)
            fsk = sin_turns(q, BITRATE * fs)
        x = x + np.where(on & (truth.gate[b] == 1), fsk, 0.0)
        # profile tone
        relt = n - truth.tone_start_sample
        ont = relt >= 0
        qt = (np.where(ont, relt, 0) * 7500) % fs
        x = x + np.where(ont, spec.tone_amp * sin_turns(qt, fs), 0.0)
        v = np.rint(x * g)
        out[s:e] = np.clip(v, -32767.0, 32767.0).astype(np.int16)
    if spec.channels > 1:
        multi = np.zeros((n_total, spec.channels), dtype=np.int16)
        multi[:, 0] = out
        for c in range(1, spec.channels):
            multi[:, c] = (hash_u64(spec.seed, 20 + c, np.arange(n_total)) & np.uint64(0xFF)).astype(np.int16) - 128
        out = multi
    if return_truth:
        return out, truth
    return out


def pcm_sha256(pcm: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(pcm).astype("<i2").tobytes()).hexdigest()


def write_wav(path: str, pcm: np.ndarray, fs: int) -> None:
    """Minimal RIFF/WAVE PCM16 writer (what scipy.io.wavfile.read expects,
    reference AXCTDprocessor.py:41)."""
    pcm = np.ascontiguousarray(pcm).astype("<i2")
    nch = 1 if pcm.ndim == 1 else pcm.shape[1]
    data = pcm.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, nch, fs, fs * nch * 2, nch * 2, 16))
        f.write(b"data" + struct.pack("<I", len(data)))
        f.write(data)


# Sample formats other than 16-bit PCM (scipy.io.wavfile.read, reference AXCTDprocessor.py:41, returns int32 for 24 /
# 32-bit PCM -- 24-bit left-justified -- and float32 / float64 for IEEE files)
WIDE_FORMATS = ("pcm24", "pcm32", "float32", "float64")


def widen(pcm: np.ndarray, fmt: str, seed: int = 0) -> np.ndarray:
    """The int16 drop re-quantised to a wider sample format with seeded low-order detail, so that the samples are
    not representable in 16 bits: what a 24-bit / float recorder would have stored."""
    rng = np.random.default_rng(777000 + seed)
    x = np.ascontiguousarray(pcm).astype(np.int64)
    if fmt == "pcm24":
        return (x * 256 + rng.integers(-128, 128, size=x.shape)).astype(np.int32)          # 24-bit values
    if fmt == "pcm32":
        return (x * 65536 + rng.integers(-32768, 32768, size=x.shape)).astype(np.int32)
    if fmt == "float32":
        return ((x + rng.uniform(-0.5, 0.5, size=x.shape)) / 32768.0).astype(np.float32)
    if fmt == "float64":
        return (x + rng.uniform(-0.5, 0.5, size=x.shape)) / 32768.0
    raise KeyError(fmt)


def write_wav_wide(path: str, samples: np.ndarray, fs: int, fmt: str) -> None:
    """RIFF/WAVE writer for widen()'s formats: 24-bit PCM (three bytes per sample), 32-bit PCM, IEEE float 32 / 64."""
    a = np.ascontiguousarray(samples)
    nch = 1 if a.ndim == 1 else a.shape[1]
    if fmt == "pcm24":
        b = a.astype("<i4").reshape(-1, 1).view(np.uint8).reshape(-1, 4)[:, :3].tobytes()
        tag, bits = 1, 24
    elif fmt == "pcm32":
        b, tag, bits = a.astype("<i4").tobytes(), 1, 32
    elif fmt == "float32":
        b, tag, bits = a.astype("<f4").tobytes(), 3, 32
    elif fmt == "float64":
        b, tag, bits = a.astype("<f8").tobytes(), 3, 64
    else:
        raise KeyError(fmt)
    bps = bits // 8
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(b)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, nch, fs, fs * nch * bps, nch * bps, bits))
        f.write(b"data" + struct.pack("<I", len(b)))
        f.write(b)


# Named workloads (BASELINE.json configs)
def config_spec(name: str, seed: int = 1) -> DropSpec:
    if name == "config1":      # 44.1 kHz, 12 min, 40 dB, default lowpass
        return DropSpec(fs=44100, duration_s=720.0, seed=seed, snr_db=40.0)
    if name == "config2":      # same drop, 10 dB SNR (bandpass / CRC-failure path)
        return DropSpec(fs=44100, duration_s=720.0, seed=seed, snr_db=10.0)
    if name == "batch44":
        return DropSpec(fs=44100, duration_s=720.0, seed=seed, snr_db=25.0)
    if name == "batch48":
        return DropSpec(fs=48000, duration_s=720.0, seed=seed, snr_db=25.0)
    raise KeyError(name)
