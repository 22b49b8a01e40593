"""svin_b200 — B200-native engine for SVIn's two data-parallel hot paths.

The compute lives in ``svin_b200/csrc`` (hand-written CUDA for sm_100a) behind the C ABI
declared in ``include/svin_b200.h``; this package only holds the ctypes view of that ABI,
host-side containers mirroring the reference's data model, and synthetic-input generators.
"""
__version__ = "0.1.0"
