set -x
PXZ_VARIANT_LEVELS=1 bash tools/variant_times.sh > gpurun_out/s2_var7.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest7.txt 2>&1; echo rc=$?
