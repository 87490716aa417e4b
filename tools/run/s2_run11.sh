set -x
PXZ_VARIANT_LEVELS=1 bash tools/variant_times.sh > gpurun_out/s2_var11.txt 2>&1
