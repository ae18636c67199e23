"""Incremental front end for a live receiver.

The reference decodes a finished WAV file, but its processor is written as a loop over 2 s iterations with a
``keepgoing`` flag and per-iteration result lists (AXCTDprocessor.py:119, :283, :612) so that a receiver can
show a profile while the probe is still falling.  ``StreamingDecoder`` gives that use case the same shape on
top of the batch engine: PCM is pushed as it arrives, ``poll()`` decodes everything received so far and
returns the profile rows that are new since the previous poll, ``finish()`` returns the decode of the complete
recording -- identical to the batch result, because it *is* the batch decode of the same samples.

Intermediate polls are provisional by nature, in the reference as here: ``readAXCTDwavfile`` normalises with
the mean and the peak of the *whole* recording (AXCTDprocessor.py:55-57), so the decode of a prefix is the
reference's decode of that prefix saved as its own file, not a prefix of the final decode.  A 12-minute drop
decodes in a few milliseconds on the device, so re-decoding the prefix once a second costs well under 1 % of
one GPU.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


class StreamingDecoder:
    def __init__(self, fs: float, settings=None, triggerrange=None, device: int = 0, engine=None,
                 min_new_seconds: float = 1.0, decimate: int = 1):
        self._own = engine is None
        self.eng = engine if engine is not None else _engine.Engine(device)
        self.fs = float(fs)
        self.cfg = self.eng.config(fs, settings=settings, triggerrange=triggerrange, decimate=decimate)
        self.keepgoing = True                   # cleared by stop(): further pushes are ignored (AXCTDprocessor.py:283)
        self.min_new = int(min_new_seconds * fs * decimate)
        self._parts, self._n = [], 0
        self._decoded_n = 0                     # samples covered by the last decode
        self._reported = 0                      # rows handed out by poll() so far
        self.last = None                        # DropResult of the last decode

    def close(self):
        if self._own and self.eng is not None:
            self.eng.close()
        self.eng = None

    def stop(self):
        self.keepgoing = False

    @property
    def n_samples(self) -> int:
        return self._n

    def push(self, pcm) -> int:
        """Append mono int16 samples; returns the number of samples held."""
        if not self.keepgoing:
            return self._n
        a = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        if a.size:
            self._parts.append(a)
            self._n += a.size
        return self._n

    def _pcm(self) -> np.ndarray:
        if len(self._parts) > 1:
            self._parts = [np.concatenate(self._parts)]
        return self._parts[0] if self._parts else np.zeros(0, dtype=np.int16)

    def _decode(self):
        pcm = self._pcm()
        self.last = self.eng.process([pcm], [self.cfg])[0]
        self._decoded_n = pcm.size
        return self.last

    def poll(self):
        """Decode what has arrived (if at least ``min_new_seconds`` are new) and return the kept profile rows
        beyond those already reported, as the structured table of ``DropResult.table()``; None if there is
        nothing new or the recording so far cannot be decoded yet (no pulse, headers incomplete ...)."""
        if self._n - self._decoded_n < max(self.min_new, 1):
            return None
        r = self._decode()
        if r.status != 0:
            return None
        tab = r.table()
        kept = tab[tab["keep"] == 1]
        new = kept[self._reported:]
        self._reported = max(self._reported, len(kept))
        return new if len(new) else None

    def finish(self):
        """Decode of the complete recording (the batch result for the same samples)."""
        self.keepgoing = False
        if self.last is None or self._decoded_n != self._n:
            self._decode()
        return self.last
