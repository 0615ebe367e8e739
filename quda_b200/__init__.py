"""quda_b200 -- B200-native Wilson / Wilson-clover Dslash engine behind QUDA's Dslash interface.

Package contents (only what the hot path needs):
  csrc/       hand-written sm_100a CUDA kernels + the C ABI (include/b200_dslash.h) -> libquda_b200.so
  lib.py      ctypes binding of the C ABI
  dslash.py   mirror of the reference entry points ApplyWilson / ApplyWilsonClover /
              ApplyWilsonCloverPreconditioned / ApplyClover / PackGhost
  fields.py   host <-> native field marshaling (FloatN orders, recon 18/12/8, block-float half, clover)
"""
from . import fields, lib  # noqa: F401
from .dslash import (ApplyClover, ApplyWilson, ApplyWilsonClover, ApplyWilsonCloverPreconditioned,  # noqa: F401
                     CloverField, ColorSpinorField, GaugeField, Halo, PackGhost)
from .lib import B200Error  # noqa: F401
