set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/s2g_bench_n1.json 2> gpurun_out/s2g_bench_n1.err; echo rc=$?
python bench.py --steps 2 --warmup 3 --batch 4 --skip-extras --issue direct > gpurun_out/s2g_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2g_launches.csv python bench.py --steps 2 --warmup 3 --batch 4 --skip-extras --issue direct > gpurun_out/s2g_ncu_launch.log 2>&1
python tools/prof_driver.py 3 > gpurun_out/s2g_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 6 -c 12 -o gpurun_out/s2g_full -f python tools/prof_driver.py 3 > gpurun_out/s2g_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2g_smoke.txt 2>&1; echo rc=$?
