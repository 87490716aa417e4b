set -x
bash tools/variant_times.sh > gpurun_out/s2_var10.txt 2>&1
bash tools/variant_times.sh >> gpurun_out/s2_var10.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest10.txt 2>&1; echo rc=$?
