set -x
bash tools/variant_times.sh > gpurun_out/s2_dup.txt 2>&1
python tools/class_times.py > gpurun_out/s2_class_times_dup.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest3.txt 2>&1; echo rc=$?
