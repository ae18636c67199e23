"""Helpers shared by the golden-vector tests (fixtures come from the unmodified
reference via oracle/make_golden.py)."""
import json
import os

import numpy as np

import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_CASES = ["g44_40db", "g44_10db", "g48_25db", "g44_stereo", "g44_bandpass", "g44_wired",
               "g44_chunk4", "g44_nopulse", "g44_chunk05", "g48_chunk8", "g44_marksp", "g48_marksp_tx",
               "g44_timeout", "g44_timeout_notone"]
DECIM_CASES = ["g96_decim"]
WIDE_CASES = ["g44_pcm24", "g48_float32", "g96_pcm24_decim"]      # 24-bit PCM / IEEE float WAV files
FULL_CASES = ["config1_720s", "config2_720s", "config5_1800s"]

_pcm_cache = {}


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.z = z
        self.spec = synth.DropSpec(**self.meta["spec"])
        self.user_settings = self.meta["user_settings"]
        self.triggerrange = self.meta["triggerrange"]

    def pcm(self):
        if self.name not in _pcm_cache:
            p = synth.generate_drop(self.spec)
            assert synth.pcm_sha256(p) == self.meta["pcm_sha256"], \
                "synthetic generator is not reproducing the PCM the golden was made from"
            _pcm_cache.clear()
            _pcm_cache[self.name] = p
        return _pcm_cache[self.name]

    def wide(self):
        """(samples, format) of a WIDE_CASES fixture's WAV file: the drop re-quantised as oracle/make_golden.py did."""
        fmt = self.meta["wav_format"]
        return synth.widen(self.pcm(), fmt, self.spec.seed), fmt

    @property
    def bits(self):
        return np.unpackbits(self.z["bits_packed"])[: self.meta["n_bits"]]

    @property
    def edges(self):
        if self.meta["n_edges"] == 0:
            return np.zeros(0, dtype=np.int64)
        return np.concatenate([self.z["edges_first"], self.z["edges_first"][0] + np.cumsum(self.z["edges_delta"].astype(np.int64))])

    @property
    def hexframes(self):
        return ["%08x" % v for v in self.z["hexframes"]]

    def trace(self):
        keys = ("s", "e", "status", "n_power", "nbits", "first_edge", "last_edge", "nrows", "nhex", "profstart")
        return [dict(zip(keys, row.tolist())) for row in self.z["trace"]]
