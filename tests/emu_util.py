"""Build / load the TEST-ONLY host emulation of the engine (see tests/emu/README.md)."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "axctdprocessor_b200", "csrc")
OUT_DIR = os.path.join(ROOT, "tests", "emu", "_build")
OUT = os.path.join(OUT_DIR, "libaxctd_emu.so")
_lib_cache = None


def build_emu():
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "axctd.h")]
    newest = max(os.path.getmtime(p) for p in srcs)
    if not os.path.isfile(OUT) or os.path.getmtime(OUT) < newest:
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-DAXCTD_EMU",
                        "-x", "c++", os.path.join(CSRC, "ax_engine.cu"), "-o", OUT], check=True)
    return OUT


def emu_engine(**options):
    global _lib_cache
    from axctdprocessor_b200 import _lib, engine
    if _lib_cache is None:
        _lib_cache = _lib.bind(C.CDLL(build_emu()))
    eng = engine.Engine(lib=_lib_cache, allow_emulation=True)
    for k, v in options.items():
        eng.set_option(k, v)
    return eng
