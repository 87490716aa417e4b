set -x
bash tools/variant_times.sh > gpurun_out/s2_twoend.txt 2>&1
python tools/class_times.py > gpurun_out/s2_class_times_twoend.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest2.txt 2>&1; echo rc=$?
python bench.py --steps 20 --warmup 5 --skip-extras > gpurun_out/s2_b_twoend.json 2> gpurun_out/s2_b_twoend.err
python bench.py --steps 20 --warmup 5 --skip-extras --streams 6 --batch 18 > gpurun_out/s2_b_twoend_s6.json 2>> gpurun_out/s2_b_twoend.err
python bench.py --steps 20 --warmup 5 --skip-extras --streams 8 > gpurun_out/s2_b_twoend_s8.json 2>> gpurun_out/s2_b_twoend.err
