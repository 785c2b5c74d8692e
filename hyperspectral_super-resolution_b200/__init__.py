"""hsr_b200 — B200-native EMIT -> Sentinel-2 pair-synthesis hot path.

GLT orthorectification, SRF band synthesis and per-band polynomial colour matching of
martasumyk/hyperspectral_super-resolution as hand-written sm_100a CUDA kernels behind a
C ABI (``include/hsr_b200.h``), with the reference's Python call surface on top:

    hsr_b200.EMIT_data.emit_proj / emit_tools   (GLT ortho;   reference EMIT_data/*.py)
    hsr_b200.s2_emit.srf / synth                (SRF bands;   reference s2_emit/srf.py, synth.py)
    hsr_b200.s2_emit.poly_regression            (poly match;  reference s2_emit/poly_regression.py)
    hsr_b200.pipeline                           (the fused ortho + SRF + polyfit + apply pass)
    hsr_b200.dist                               (granule / tile sharding, moment all-reduce)

There is no CPU fallback: every compute entry point raises if the CUDA library or a GPU is
missing.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (ctypes binding of libhsr_b200.so; lazy-loads the .so)
