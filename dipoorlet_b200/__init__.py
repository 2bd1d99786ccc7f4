"""dipoorlet_b200 — B200-native (sm_100a) implementation of Dipoorlet's activation-
calibration and rounding-finetune hot path behind the reference's plugin API.

See DESIGN.md for the path, its boundary and the kernels; include/dpl_b200.h for the
C-ABI; INTEGRATION.md for how the reference binds to it.
"""
__version__ = "0.1.0"
