#!/bin/bash
# times every pixlzr-rust_b200/var_*.so (and the default build) on the bench frame and per level: run under gpurun
for so in pixlzr-rust_b200/libpixlzr_b200.so pixlzr-rust_b200/var_*.so; do
  [ -f "$so" ] || continue
  PXZ_LIB=$PWD/$so timeout 120 python tools/kernel_times.py 10
  if [ -n "$PXZ_VARIANT_LEVELS" ]; then echo "-- $so"; PXZ_LIB=$PWD/$so timeout 180 python tools/class_times.py; fi
done
