#!/bin/bash
# Build the library variants the round-2 experiments compare (run HERE, before the gpurun call; each .so is ~80 MB, so
# delete them again afterwards: `rm quda_b200/libquda_b200_*.so; rm -rf quda_b200/csrc/_obj_*`):
#   _i2f3   half precision: 7/8 of the int16->fp32 conversions on the conversion unit (B2_I2F_NATIVE=3; 1824 vs 2208
#           SASS instructions per site for half recon-12, tools/sass_mix.py)
#   _i2f2   all conversions on the conversion unit
#   _nodef  half precision with scale-on-load arithmetic (the pre-deferred-scaling kernels, 2416 instructions)
set -e
cd "$(dirname "$0")/../quda_b200/csrc"
make -j"$(nproc)" VARIANT=_i2f3 VFLAGS="-DB2_I2F_NATIVE=3" > /dev/null
make -j"$(nproc)" VARIANT=_i2f2 VFLAGS="-DB2_I2F_NATIVE=2" > /dev/null
make -j"$(nproc)" VARIANT=_nodef VFLAGS="-DB2_HALF_DEFERRED_SCALE=0" > /dev/null
ls -la ../libquda_b200*.so
