set -x
PXZ_VARIANT_LEVELS=1 bash tools/variant_times.sh > gpurun_out/s2_var5.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest5.txt 2>&1; echo rc=$?
