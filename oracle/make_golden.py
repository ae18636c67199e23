"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  For every case in
CASES it regenerates the synthetic drop from its seed (synth.py), runs the
reference's own classes through oracle/ref_shim.py and stores the reference's
outputs and per-chunk intermediates as a compressed fixture.  The fixtures (not
the reference) travel to the GPU box.

usage:  python oracle/make_golden.py [case ...]
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_shim  # noqa: E402
import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (DropSpec kwargs, reference user_settings (internal key names), triggerrange override, full detail?)
CASES = {
    "g44_40db":   (dict(fs=44100, duration_s=60.0, seed=1, snr_db=40.0), {}, None, True),
    "g44_10db":   (dict(fs=44100, duration_s=76.0, seed=2, snr_db=10.0), {}, None, True),
    "g48_25db":   (dict(fs=48000, duration_s=60.0, seed=3, snr_db=25.0), {}, None, True),
    "g44_stereo": (dict(fs=44100, duration_s=52.0, seed=4, snr_db=30.0, channels=2), {}, None, True),
    "g44_bandpass": (dict(fs=44100, duration_s=60.0, seed=5, snr_db=40.0), {"usebandpass": True}, None, True),
    "g44_wired":  (dict(fs=44100, duration_s=60.0, seed=6, snr_db=20.0),
                   {"minr400": 1.5, "mindr7500": 1.0, "deadfreq": 2500.0, "refreshrate": 1.0,
                    "mark_space_freqs": [400.0, 800.0]}, [31, -1], True),
    "g44_chunk4": (dict(fs=44100, duration_s=60.0, seed=7, snr_db=15.0), {"refreshrate": 4.0}, None, True),
    "g96_decim":  (dict(fs=96000, duration_s=56.0, seed=8, snr_db=30.0), {}, None, True),
    "g44_nopulse": (dict(fs=44100, duration_s=12.0, seed=9, snr_db=30.0, lead_in_s=30.0), {}, None, True),
    # chunk sizes at both ends of BASELINE config 5's sweep (-l 0.5 .. 8 x fs), wired through 'refreshrate'
    "g44_chunk05": (dict(fs=44100, duration_s=60.0, seed=10, snr_db=20.0), {"refreshrate": 0.5}, None, True),
    "g48_chunk8": (dict(fs=48000, duration_s=76.0, seed=11, snr_db=15.0), {"refreshrate": 8.0}, None, True),
    # custom mark / space (-m / -n): detuned bins on a nominal signal, and a probe that transmits 420 / 780 Hz
    "g44_marksp": (dict(fs=44100, duration_s=60.0, seed=12, snr_db=25.0),
                   {"mark_space_freqs": [410.0, 790.0], "deadfreq": 2800.0}, None, True),
    "g48_marksp_tx": (dict(fs=48000, duration_s=60.0, seed=13, snr_db=30.0, mark_hz=420, space_hz=780),
                      {"mark_space_freqs": [420.0, 780.0]}, None, True),
    # latest trigger time (-b): the elif of AXCTDprocessor.py:404-408 only runs once status is already 2 (or while
    # mean7500pwr is NaN); it then overwrites profstartind while firstpointtime keeps the old value
    "g44_timeout": (dict(fs=44100, duration_s=64.0, seed=14, snr_db=25.0), {}, [30, 41], True),
    "g44_timeout_notone": (dict(fs=44100, duration_s=60.0, seed=15, snr_db=25.0, tone_amp=0.0), {}, [30, 36], True),
    # sample formats other than 16-bit PCM (scipy reads them too, AXCTDprocessor.py:41): WIDE below names the format
    "g44_pcm24": (dict(fs=44100, duration_s=52.0, seed=16, snr_db=20.0), {}, None, True),
    "g48_float32": (dict(fs=48000, duration_s=52.0, seed=17, snr_db=30.0), {}, None, True),
    "g96_pcm24_decim": (dict(fs=96000, duration_s=52.0, seed=18, snr_db=25.0), {}, None, True),
    "config1_720s": (dict(fs=44100, duration_s=720.0, seed=1, snr_db=40.0), {}, None, False),
    "config2_720s": (dict(fs=44100, duration_s=720.0, seed=1, snr_db=10.0), {}, None, False),
    # BASELINE config 5 stand-in: a 30-minute recording at 8 dB SNR decoded with a swept parameter point
    # (chunk 4 x fs, dead frequency 2500 Hz, detuned mark / space)
    "config5_1800s": (dict(fs=44100, duration_s=1800.0, seed=5, snr_db=8.0),
                      {"refreshrate": 4.0, "deadfreq": 2500.0, "mark_space_freqs": [405.0, 795.0]}, None, False),
}


# cases whose WAV file holds synth.widen(pcm, format, seed) instead of the int16 drop
WIDE = {"g44_pcm24": "pcm24", "g48_float32": "float32", "g96_pcm24_decim": "pcm24"}


def _sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _jsonable(o):
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    return o


def make_case(name: str) -> None:
    spec_kw, user_settings, trig, full = CASES[name]
    spec = synth.DropSpec(**spec_kw)
    pcm = synth.generate_drop(spec)
    with tempfile.TemporaryDirectory() as td:
        wav = os.path.join(td, name + ".wav")
        if name in WIDE:
            synth.write_wav_wide(wav, synth.widen(pcm, WIDE[name], spec.seed), spec.fs, WIDE[name])
        else:
            synth.write_wav(wav, pcm, spec.fs)
        t0 = time.time()
        err = None
        try:
            ap, trace = ref_shim.run_processor(wav, user_settings=user_settings, triggerrange=trig)
        except Exception as ex:  # the reference's own crash is part of the contract
            err = f"{type(ex).__name__}: {ex}"
            ap, trace = None, []
        dt = time.time() - t0
    meta = dict(name=name, spec=spec_kw, user_settings=user_settings, triggerrange=trig,
                pcm_sha256=synth.pcm_sha256(pcm), reference_seconds=dt, error=err,
                generator="oracle/make_golden.py", numpy=np.__version__, wav_format=WIDE.get(name))
    arrays = {}
    if ap is not None:
        edges = ap._all_edges
        meta.update(
            f_s=float(ap.f_s), numpoints=int(ap.numpoints), firstpulse400=int(ap.firstpulse400),
            profstartind=int(ap.profstartind), high_bit_scale=float(ap.high_bit_scale),
            mean7500pwr=float(ap.mean7500pwr),
            metadata=_jsonable(ap.metadata), n_bits=int(len(ap._all_bits)), n_edges=int(len(edges)),
            bits_sha256=_sha(ap._all_bits.astype(np.uint8)), edges_sha256=_sha(edges.astype("<i8")),
            hexframes_sha256=hashlib.sha256("\n".join(ap.hexframes).encode()).hexdigest(),
            n_hexframes=len(ap.hexframes), n_rows=len(ap.time),
            tcoeff=_jsonable(ap.tcoeff), ccoeff=_jsonable(ap.ccoeff), zcoeff=_jsonable(ap.zcoeff))
        arrays["bits_packed"] = np.packbits(ap._all_bits.astype(np.uint8))
        arrays["edges_first"] = edges[:1].astype(np.int64)
        arrays["edges_delta"] = np.diff(edges).astype(np.int32)
        arrays["trace"] = np.array([[r.get(k, -1) for k in ("s", "e", "status", "n_power", "nbits", "first_edge",
                                                            "last_edge", "nrows", "nhex", "profstart")] for r in trace], dtype=np.int64)
        # every CRC-valid profile frame BEFORE rounding and QC (parse.py:92): the 1e-6 bar is held on these
        for k, v in ap._raw.items():
            arrays["raw_" + k] = v
        arrays["trace_scale"] = np.array([r.get("scale") or 0.0 for r in trace], dtype=np.float64)
        for k in ("time", "depth", "temperature", "conductivity", "salinity", "r400_prof", "r7500_prof"):
            arrays[k] = np.asarray(getattr(ap, k), dtype=np.float64)
        arrays["hexframes"] = np.array([int(h, 16) for h in ap.hexframes], dtype=np.uint32)
        arrays["r400"] = np.asarray(ap.r400, dtype=np.float64) if full else np.asarray(ap.r400, dtype=np.float32)
        arrays["r7500"] = np.asarray(ap.r7500, dtype=np.float64) if full else np.asarray(ap.r7500, dtype=np.float32)
        arrays["power_inds"] = np.asarray(ap.power_inds, dtype=np.int64)
        if full:
            arrays["conf"] = ap._all_conf
        # output file exactly as the reference CLI writes it (processAXCTD.py:143-183)
        from axctd_oracle import format_output
        settings_cli = {"minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0, "pointsperloop": 100000,
                        "triggerrange": [30, -1], "mark_space_freqs": [400.0, 800.0], "use_bandpass": False}
        try:
            meta["output_text"] = format_output(ap, name + ".wav", [0, -1], settings_cli)
        except KeyError as ex:
            meta["output_text_error"] = f"KeyError: {ex}"
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name}: ref {dt:.1f}s err={err} rows={meta.get('n_rows')} bits={meta.get('n_bits')} "
          f"first={meta.get('firstpulse400')} prof={meta.get('profstartind')} scale={meta.get('high_bit_scale')}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        make_case(n)
