set -x
bash tools/variant_times.sh > gpurun_out/s2_var9.txt 2>&1
PXZ_LIB=$PWD/pixlzr-rust_b200/var_exp3.so python tools/class_times.py >> gpurun_out/s2_var9.txt 2>&1
