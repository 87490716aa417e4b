python bench.py --steps 20 --warmup 5 > gpurun_out/s2_last_bench.json 2> gpurun_out/s2_last_bench.err; echo rc=$?
