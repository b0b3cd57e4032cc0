"""
timbre_trap_b200 - the data-parallel hot path of sony/timbre-trap (NSGT CQT analysis,
SoundStream-style conv encoder/decoder, CQT synthesis, the loss step) on hand-written
sm_100a CUDA kernels behind a C ABI (include/timbre_trap_b200.h).

`timbre_trap_b200.framework` mirrors `timbre_trap.framework` of the reference.
"""

__version__ = '0.1.0'
