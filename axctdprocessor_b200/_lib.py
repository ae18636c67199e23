"""ctypes binding of the C ABI declared in include/axctd.h.

The product has exactly one compute path: the CUDA library built in-tree by
``__graft_entry__.build()`` (axctdprocessor_b200/libaxctd.so).  If it is
missing, or no CUDA device can be opened, loading fails loudly; there is no CPU
fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libaxctd.so")

MAX_SECTIONS = 6
ABI_VERSION = 4          # AXCTD_ABI_VERSION of include/axctd.h


class ConfigDesc(C.Structure):
    _fields_ = [
        ("fs", C.c_double),
        ("n_power", C.c_int32), ("d_pcm", C.c_int32), ("npcm", C.c_int32), ("chunk_len", C.c_int32),
        ("pad", C.c_int32), ("bit_inset", C.c_int32), ("bitrate", C.c_int32), ("n_sections", C.c_int32),
        ("sos", (C.c_double * 6) * MAX_SECTIONS),
        ("max_pole_radius", C.c_double),
        ("bit_cs", C.POINTER(C.c_double)), ("bit_cs_len", C.c_int32), ("reserved0", C.c_int32),
        ("tone_cs", C.POINTER(C.c_double)),
        ("min_r400", C.c_double), ("min_dr7500", C.c_double),
        ("trigger_from_s", C.c_double), ("trigger_to_s", C.c_double), ("high_bit_scale0", C.c_double),
        ("zcoeff", C.c_double * 4), ("tcoeff", C.c_double * 4), ("ccoeff", C.c_double * 4),
        ("tlims", C.c_double * 2), ("slims", C.c_double * 2),
        ("temp_lut", C.POINTER(C.c_double)), ("lut_len", C.c_int32),
        ("hist_edges", C.POINTER(C.c_double)), ("hist_centers", C.POINTER(C.c_double)),
        ("n_hist_edges", C.c_int32),
        ("decimate", C.c_int32), ("decim_sections", C.c_int32), ("decim_padlen", C.c_int32),
        ("decim_sos", (C.c_double * 6) * MAX_SECTIONS), ("decim_zi", (C.c_double * 2) * MAX_SECTIONS),
        ("decim_pole_radius", C.c_double),
    ]


class DropSummary(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("status_chunk", C.c_int32),
        ("numpoints", C.c_int64), ("f_s", C.c_double),
        ("firstpulse400", C.c_int64), ("profstartind", C.c_int64),
        ("firstpointtime", C.c_double), ("mean7500pwr", C.c_double), ("high_bit_scale", C.c_double),
        ("n_chunks", C.c_int32), ("first_demod_chunk", C.c_int32), ("profile_chunk", C.c_int32),
        ("header_read", C.c_int32 * 3), ("header_chunk", C.c_int32 * 3),
        ("n_bits", C.c_int64), ("n_edges", C.c_int64), ("n_power", C.c_int64),
        ("n_frames", C.c_int64), ("n_rows", C.c_int64), ("n_hex", C.c_int64), ("n_crossings", C.c_int64),
        ("n_uncertain", C.c_int32), ("n_chain_fixups", C.c_int32),
        ("n_guard_hits", C.c_int32), ("n_guard_confirmed", C.c_int32),
        ("pcm_sum", C.c_int64), ("pcm_ampl", C.c_int32), ("n_recheck", C.c_int32),
        ("win32_max_rel_err", C.c_float), ("n_frame_respec", C.c_int32),
        ("frame_data", (C.c_uint16 * 72) * 2), ("counter_found", (C.c_uint8 * 72) * 2),
        ("header_parsed", C.c_int32 * 2),
        ("zcoeff", C.c_double * 4), ("tcoeff", C.c_double * 4), ("ccoeff", C.c_double * 4),
        ("zcoeff_valid", C.c_int32 * 4), ("tcoeff_valid", C.c_int32 * 4), ("ccoeff_valid", C.c_int32 * 4),
        ("zcoeff_used", C.c_double * 4), ("tcoeff_used", C.c_double * 4), ("ccoeff_used", C.c_double * 4),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("edge_index", C.c_int64), ("word", C.c_uint32), ("chunk", C.c_int32),
        ("cint", C.c_int32), ("tint", C.c_int32), ("keep", C.c_int32), ("hex_returned", C.c_int32),
        ("time_s", C.c_double), ("depth", C.c_double), ("temperature", C.c_double),
        ("conductivity", C.c_double), ("salinity", C.c_double), ("r400", C.c_double), ("r7500", C.c_double),
        ("time_raw", C.c_double), ("depth_raw", C.c_double), ("temperature_raw", C.c_double),
        ("conductivity_raw", C.c_double), ("salinity_raw", C.c_double), ("r400_raw", C.c_double),
        ("r7500_raw", C.c_double),
    ]


class Chunk(C.Structure):
    _fields_ = [
        ("s", C.c_int64), ("e", C.c_int64), ("status", C.c_int32), ("n_power_total", C.c_int32),
        ("n_bits", C.c_int32), ("first_edge", C.c_int32), ("last_edge", C.c_int32), ("n_head_edges", C.c_int32),
        ("n_rows", C.c_int32), ("n_hex", C.c_int32), ("scale", C.c_double), ("profstartind", C.c_int64),
    ]


class Row(C.Structure):
    _fields_ = [
        ("word", C.c_uint32), ("time_c", C.c_int32), ("depth_c", C.c_int32),
        ("temperature_c", C.c_int16), ("conductivity_c", C.c_int16), ("salinity_c", C.c_int16),
        ("r400_c", C.c_int16), ("r7500_c", C.c_int16), ("flags", C.c_uint16),
    ]


ROW_KEEP, ROW_HEX, ROW_WIDE, ROW_NAN, ROW_NAN16 = 1, 2, 4, -2147483648, -32768


class SynthDesc(C.Structure):
    _fields_ = [
        ("n_total", C.c_int64), ("n0", C.c_int64), ("tone_start", C.c_int64), ("fs", C.c_int64),
        ("key1", C.c_uint64), ("key2", C.c_uint64),
        ("nscale", C.c_double), ("gain", C.c_double), ("tone_amp", C.c_double),
        ("sin_coef", C.c_double * 9),
        ("bits", C.c_void_p), ("gate", C.c_void_p), ("parity", C.c_void_p), ("nslots", C.c_int64),
    ]


SYMBOLS = [
    "axctd_abi_version", "axctd_has_cuda", "axctd_struct_size", "axctd_engine_create", "axctd_engine_destroy", "axctd_last_error",
    "axctd_engine_set_option", "axctd_engine_set_stream", "axctd_engine_launch_count", "axctd_config_create", "axctd_batch_create",
    "axctd_batch_destroy", "axctd_batch_upload", "axctd_batch_upload_interleaved", "axctd_batch_upload_f64", "axctd_batch_copy_from", "axctd_batch_device_pcm", "axctd_batch_run",
    "axctd_batch_run_async", "axctd_batch_finish", "axctd_batch_timing", "axctd_batch_phase_ms", "axctd_batch_summary",
    "axctd_batch_rows", "axctd_batch_frames", "axctd_batch_chunks", "axctd_batch_bits", "axctd_batch_edges", "axctd_batch_power",
    "axctd_synth_fill", "axctd_batch_download", "axctd_calib_eval",
    "axctd_batch_stream_begin", "axctd_batch_stream_append", "axctd_batch_stream_run",
]


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argument / result types to every symbol of include/axctd.h."""
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    P = C.POINTER
    sig = {
        "axctd_abi_version": (i32, []),
        "axctd_has_cuda": (i32, []),
        "axctd_struct_size": (i32, [i32]),
        "axctd_engine_create": (i32, [i32, P(vp)]),
        "axctd_engine_destroy": (None, [vp]),
        "axctd_last_error": (C.c_char_p, [vp]),
        "axctd_engine_set_option": (i32, [vp, C.c_char_p, dbl]),
        "axctd_engine_launch_count": (i64, [vp]),
        "axctd_engine_set_stream": (i32, [vp, vp]),
        "axctd_config_create": (i32, [vp, P(ConfigDesc), P(i32)]),
        "axctd_batch_create": (i32, [vp, i32, P(i64), P(C.c_int32), P(vp)]),
        "axctd_batch_destroy": (None, [vp]),
        "axctd_batch_upload": (i32, [vp, i32, vp, i64]),
        "axctd_batch_upload_interleaved": (i32, [vp, i32, vp, i64, i32]),
        "axctd_batch_upload_f64": (i32, [vp, i32, vp, i64]),
        "axctd_batch_copy_from": (i32, [vp, i32, vp, i32, i64, i64]),
        "axctd_batch_device_pcm": (i32, [vp, i32, P(vp)]),
        "axctd_batch_run": (i32, [vp]),
        "axctd_batch_run_async": (i32, [vp]),
        "axctd_batch_finish": (i32, [vp]),
        "axctd_batch_timing": (i32, [vp, P(dbl), P(dbl), P(dbl)]),
        "axctd_batch_phase_ms": (i32, [vp, P(dbl)]),
        "axctd_batch_summary": (i32, [vp, i32, P(DropSummary)]),
        "axctd_batch_rows": (i64, [vp, i32, vp, i64]),
        "axctd_batch_frames": (i64, [vp, i32, vp, i64]),
        "axctd_batch_chunks": (i64, [vp, i32, vp, i64]),
        "axctd_batch_bits": (i64, [vp, i32, vp, vp, i64]),
        "axctd_batch_edges": (i64, [vp, i32, vp, vp, vp, i64]),
        "axctd_batch_power": (i64, [vp, i32, vp, vp, vp, i64]),
        "axctd_synth_fill": (i32, [vp, i32, P(SynthDesc)]),
        "axctd_batch_download": (i32, [vp, i32, vp, i64]),
        "axctd_calib_eval": (i32, [vp, vp, vp, vp, i32, vp, vp, vp]),
        "axctd_batch_stream_begin": (i32, [vp, P(dbl), P(dbl)]),
        "axctd_batch_stream_append": (i32, [vp, i32, vp, i64]),
        "axctd_batch_stream_run": (i32, [vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, st in enumerate((ConfigDesc, DropSummary, Frame, Chunk, Row)):
        if lib.axctd_struct_size(which) != C.sizeof(st):
            raise ImportError(f"ABI struct size mismatch for {st.__name__}: "
                              f"{lib.axctd_struct_size(which)} != {C.sizeof(st)}")
    return lib


_cached = None


def load() -> C.CDLL:
    """The in-tree CUDA library; raises if it has not been built."""
    global _cached
    if _cached is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()').  axctdprocessor_b200 has no CPU fallback.")
        lib = bind(C.CDLL(LIB_PATH))
        if lib.axctd_abi_version() != ABI_VERSION:
            raise ImportError("libaxctd.so ABI version mismatch")
        if not lib.axctd_has_cuda():
            raise ImportError("libaxctd.so was not built with CUDA kernels")
        _cached = lib
    return _cached
