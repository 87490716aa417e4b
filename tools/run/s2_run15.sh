set -x
PXZ_VARIANT_LEVELS=1 bash tools/variant_times.sh > gpurun_out/s2_var15.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest15.txt 2>&1; echo rc=$?
timeout 500 python tests/tools/fuzz_parity.py 3000 91 > gpurun_out/s2_fuzz15.txt 2>&1; echo rc=$?
