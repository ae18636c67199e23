"""Incremental decoder for a live receiver.

The reference decodes a finished WAV file, but its processor is written as a loop over 2 s iterations with a
``keepgoing`` flag and per-iteration result lists (AXCTDprocessor.py:119, :283-338, :612) so that a receiver can
show a profile while the probe is still falling.  ``StreamingDecoder`` runs that loop on the device as the audio
arrives (C ABI: axctd_batch_stream_begin / _append / _run):

* ``push`` copies the new samples -- and only those -- to the device, behind the ones already there;
* ``poll`` decodes the iterations that have become complete since the last poll on top of the state the device
  keeps (tone block sums, crossing records, chunk chain, level history, bit buffer, header and calibration state) and
  returns the profile rows of exactly those iterations.  Rows handed out are final: ``finish`` repeats them unchanged;
* ``finish`` closes the recording (the end-of-file rules of AXCTDprocessor.py:295-300 for what is left) and returns
  the complete result.

One thing a live decoder cannot do as the reference does: ``readAXCTDwavfile`` normalises with the mean and the peak
of the WHOLE file (AXCTDprocessor.py:55-57).  The decoder fixes the two numbers up front -- ``norm=(dc, ampl)`` from
the caller, or the reference's formula over the first ``norm_seconds`` of audio -- and its result is the reference's
for the recording normalised with them (every output of the path is a sign, a ratio or a calibrated frame integer,
so the peak only scales intermediate values).  Recordings above 50 kHz are not streamed (the reference halves them
with a forward-backward filter over the whole file, AXCTDprocessor.py:60-62).
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


def prefix_normalisation(pcm) -> tuple:
    """(dc, ampl) as readAXCTDwavfile takes them (AXCTDprocessor.py:55-56), over the samples given."""
    x = np.asarray(pcm, dtype=np.int16).reshape(-1)
    return float(np.mean(x)), float(np.max(np.abs(x)))


class StreamingDecoder:
    def __init__(self, fs: float, settings=None, triggerrange=None, device: int = 0, engine=None,
                 max_seconds: float = 1800.0, norm=None, norm_seconds: float = 2.0, min_new_seconds: float = 0.0):
        if fs > 50000:
            raise ValueError("recordings above 50 kHz cannot be streamed (AXCTDprocessor.py:60-62 filters the whole file backwards)")
        self._own = engine is None
        self.eng = engine if engine is not None else _engine.Engine(device)
        self.fs = float(fs)
        self.cfg = self.eng.config(fs, settings=settings, triggerrange=triggerrange)
        self.capacity = int(max_seconds * fs)
        self.keepgoing = True                   # cleared by stop(): further pushes are ignored (AXCTDprocessor.py:283)
        self.min_new = int(min_new_seconds * fs)
        self.norm = tuple(norm) if norm is not None else None
        self._norm_n = int(norm_seconds * fs)
        self._held = []                         # samples waiting for the normalisation to be fixed
        self._n = 0                             # samples received
        self._n_dev = 0                         # samples on the device
        self._decoded_n = 0                     # samples on the device at the last run
        self._rows_out = 0                      # table rows (frames) handed out so far
        self._chunks_out = 0                    # iterations handed out so far
        self.batch = None
        self.last = None                        # DropResult of the last run
        self.closed = False
        self.runs = []                          # per run: samples on the device, iterations decoded, device ms

    def close(self):
        if self.batch is not None:
            self.batch.close()
            self.batch = None
        if self._own and self.eng is not None:
            self.eng.close()
        self.eng = None

    def stop(self):
        self.keepgoing = False

    @property
    def n_samples(self) -> int:
        return self._n

    def _start(self):
        self.batch = self.eng.batch([self.capacity], [self.cfg])
        self.batch.stream_begin(self.norm[0], self.norm[1])

    def _append(self, a):
        if self.batch is None:
            self._start()
        self.batch.stream_append(0, a)
        self._n_dev += a.size

    def push(self, pcm) -> int:
        """Append mono int16 samples; returns the number of samples received so far."""
        if not self.keepgoing or self.closed:
            return self._n
        a = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        if a.size == 0:
            return self._n
        if self._n + a.size > self.capacity:
            raise ValueError("recording longer than max_seconds")
        self._n += a.size
        if self.norm is None:
            self._held.append(a)
            if sum(x.size for x in self._held) >= self._norm_n:
                self._fix_norm()
        else:
            self._append(a)
        return self._n

    def _fix_norm(self):
        held = np.concatenate(self._held) if self._held else np.zeros(0, dtype=np.int16)
        self._held = []
        if self.norm is None:
            if held.size == 0:
                raise ValueError("no samples to take the normalisation from")
            self.norm = prefix_normalisation(held[:self._norm_n] if self._norm_n > 0 else held)   # the first norm_seconds exactly
        if held.size:
            self._append(held)

    def _run(self, final):
        self.batch.stream_run(final)
        self._decoded_n = self._n_dev
        self.last = self.batch.result(0, full=False)
        t = self.batch.timing()
        self.runs.append(dict(samples=self._n_dev, iterations=int(self.last.summary.n_chunks), device_ms=t["total_ms"],
                              filter_ms=t["filter_ms"]))
        return self.last

    def _new_rows(self, r):
        """Rows of the iterations decoded since the last hand-out (kept rows only, as the reference's lists)."""
        tab = r.table()
        new = tab[self._rows_out:]
        self._rows_out = len(tab)
        self._chunks_out = len(r.chunks)
        return new[new["keep"] == 1]

    def poll(self):
        """Decode the iterations that have become complete (if at least ``min_new_seconds`` arrived since the last
        run) and return their kept profile rows as the structured table of ``DropResult.table()``; None when there is
        nothing new (no complete iteration yet, profile not started, normalisation not fixed yet)."""
        if self.closed or self.batch is None or self._n_dev - self._decoded_n < max(self.min_new, 1):
            return None
        r = self._run(False)
        r.raise_for_status()
        new = self._new_rows(r)
        return new if len(new) else None

    def finish(self):
        """Close the recording and return the complete DropResult (rows already handed out are in it unchanged)."""
        self.keepgoing = False
        if not self.closed:
            if self.batch is None:
                self._fix_norm()
            if self.batch is None:
                raise ValueError("no samples")
            self._run(True)
            self.closed = True
            self.last = self.batch.result(0, full=True)
        return self.last

    def finish_rows(self):
        """finish(), and the kept rows not yet handed out by poll()."""
        r = self.finish()
        return r, self._new_rows(r)
