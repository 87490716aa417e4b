set -x
bash tools/variant_times.sh > gpurun_out/s2_var4.txt 2>&1
python tools/class_times.py > gpurun_out/s2_class_times4.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest4.txt 2>&1; echo rc=$?
