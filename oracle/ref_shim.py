"""TEST INFRASTRUCTURE (oracle tooling) -- run the UNMODIFIED reference.

Imports the reference's own modules from /root/reference (read-only, present
only in the build container, never on the GPU box) under a non-invasive shim
 and captures outputs plus per-chunk intermediates.  It is
used ONLY by oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and by the container-only cross-check tests (skipped when
/root/reference is absent).  Nothing in the product package imports this.

Shim (no reference file is edited or copied):
  * numpy-2 aliases ``np.float`` / ``np.NaN`` (reference AXCTDprocessor.py:57,149; parse.py:123)
  * stub ``matplotlib`` / ``matplotlib.pyplot`` (imported, never used: processAXCTD.py:36)
  * ``gsw`` stand-in exposing ``SP_from_C`` from oracle/pss78.py (gsw is not installed)
  * cwd set to the reference directory so ``./temp_LUT.txt`` resolves (AXCTDprocessor.py:130)
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference exists only in the build container; __graft_entry__.build() installs an unmodified copy of its
# five files under baseline/_ref/ (git-ignored, shipped to the GPU box) so that bench.py --impl reference can time the
# real reference there.  Tests and fixtures are generated from /root/reference itself.
_INSTALLED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _default_ref_dir():
    env = os.environ.get("AXCTD_REFERENCE_DIR")
    if env:
        return env
    if os.path.isfile("/root/reference/AXCTDprocessor.py"):
        return "/root/reference"
    return _INSTALLED


REF_DIR = _default_ref_dir()


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "AXCTDprocessor.py"))


def _install_shim():
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        mp = types.ModuleType("matplotlib.pyplot")
        m.pyplot = mp
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = mp
    if "gsw" not in sys.modules:
        if _HERE not in sys.path:
            sys.path.insert(0, _HERE)
        import pss78
        g = types.ModuleType("gsw")
        g.SP_from_C = pss78.SP_from_C
        sys.modules["gsw"] = g
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


@contextlib.contextmanager
def _in_ref_dir():
    old = os.getcwd()
    os.chdir(REF_DIR)
    try:
        yield
    finally:
        os.chdir(old)


def load_modules():
    """Returns the reference's (AXCTDprocessor, demodulate, parse, processAXCTD) modules."""
    if not available():
        raise RuntimeError("reference tree not present")
    _install_shim()
    import importlib
    mods = []
    for name in ("AXCTDprocessor", "demodulate", "parse", "processAXCTD"):
        mods.append(importlib.import_module(name))
    return tuple(mods)


def run_cli(argv, quiet=True):
    """Faithful mode: the reference CLI exactly as shipped (processAXCTD.py:47)."""
    AX, dm, ps, cli = load_modules()
    old = sys.argv
    sys.argv = ["processAXCTD.py"] + [os.path.abspath(a) if os.path.exists(a) else a for a in argv]
    try:
        with _in_ref_dir():
            if quiet:
                with contextlib.redirect_stdout(io.StringIO()):
                    cli.main()
            else:
                cli.main()
    finally:
        sys.argv = old


def run_processor(wav, user_settings=None, triggerrange=None, capture=True, quiet=True):
    """Drive the reference class directly.  ``user_settings`` uses the
    processor's INTERNAL key names (wired mode);
    ``{}`` reproduces the faithful CLI behaviour.  Returns (ap, trace) where
    trace is a list with one dict per demodulated chunk."""
    AX, dm, ps, cli = load_modules()
    wav = os.path.abspath(wav)
    trace = []
    with _in_ref_dir():
        ap = AX.AXCTD_Processor(wav, user_settings=dict(user_settings or {}))
        if triggerrange is not None:
            ap.triggerrange = list(triggerrange)
        if capture:
            orig_demod = dm.demodulate_axctd
            orig_iter = ap.iterate_AXCTD_process
            state = {}

            def demod_wrap(pcm, *a, **k):
                bits, conf, edges, nxt = orig_demod(pcm, *a, **k)
                state["demod"] = dict(nbits=len(bits), first_edge=int(edges[0]),
                                      last_edge=int(edges[-1]), next_ind=int(nxt),
                                      scale=float(a[-1]) if a else None)
                return bits, conf, edges, nxt

            def iter_wrap(e):
                s = int(ap.demodbufferstartind)
                state.clear()
                data = orig_iter(e)
                rec = dict(s=s, e=int(e), status=int(ap.status), n_power=len(ap.power_inds),
                           nrows=(len(data[1]) if len(data) > 1 else 0),
                           nhex=(len(data[8]) if len(data) > 1 else 0), profstart=int(ap.profstartind))
                rec.update(state.get("demod", {}))
                trace.append(rec)
                return data

            dm.demodulate_axctd = demod_wrap
            ap.iterate_AXCTD_process = iter_wrap
        # keep every demodulated bit / edge for golden capture, and the UNROUNDED per-frame values
        # parse.parse_bitstream_to_profile returns (parse.py:92) before AXCTDprocessor.py:559-566 rounds them
        allbits, alledges, allconf = [], [], []
        raw = {k: [] for k in ("time", "depth", "temperature", "conductivity", "salinity", "r400", "r7500")}
        orig_parse = ps.parse_bitstream_to_profile
        if capture:
            def parse_keep(*a, **k):
                out = orig_parse(*a, **k)
                for key, vals in zip(("time", "depth", "temperature", "conductivity", "salinity", "r400", "r7500"), out[1:8]):
                    raw[key].extend(float(v) for v in vals)
                return out
            ps.parse_bitstream_to_profile = parse_keep
        if capture:
            inner = dm.demodulate_axctd

            def demod_keep(pcm, *a, **k):
                bits, conf, edges, nxt = inner(pcm, *a, **k)
                base = int(ap.demodbufferstartind)
                allbits.extend(int(b) for b in bits)
                alledges.extend(int(x) + base for x in edges)
                allconf.extend(float(c) for c in conf)
                return bits, conf, edges, nxt
            dm.demodulate_axctd = demod_keep
        try:
            if quiet:
                with contextlib.redirect_stdout(io.StringIO()):
                    ap.run()
            else:
                ap.run()
        finally:
            if capture:
                dm.demodulate_axctd = orig_demod
                ps.parse_bitstream_to_profile = orig_parse
    ap._raw = {k: np.array(v, dtype=np.float64) for k, v in raw.items()}
    ap._all_bits = np.array(allbits, dtype=np.uint8)
    ap._all_edges = np.array(alledges, dtype=np.int64)
    ap._all_conf = np.array(allconf, dtype=np.float64)
    return ap, trace
