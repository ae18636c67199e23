"""B200-native AXCTD demodulation and decoding engine.

Drop-in for the entry points of cdens/AXCTDprocessor:

    from axctdprocessor_b200 import AXCTDprocessor, processAXCTD
    ap = AXCTDprocessor.AXCTD_Processor("drop.wav"); ap.run()
    processAXCTD.main(["-i", "drop.wav", "-o", "out.txt"])

All signal processing runs in hand-written CUDA kernels (csrc/) behind the C
ABI of include/axctd.h; there is no CPU fallback.
"""
__all__ = ["AXCTDprocessor", "processAXCTD", "engine", "batch", "segment", "stream"]
