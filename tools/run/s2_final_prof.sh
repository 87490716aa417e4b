set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/s2f_bench_n1.json 2> gpurun_out/s2f_bench_n1.err; echo rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2f_bench_ref.json 2> gpurun_out/s2f_bench_ref.err; echo rc=$?
python bench.py --steps 2 --warmup 3 --batch 4 --skip-extras --issue direct > gpurun_out/s2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2f_launches.csv python bench.py --steps 2 --warmup 3 --batch 4 --skip-extras --issue direct > gpurun_out/s2f_ncu_launch.log 2>&1
python tools/prof_driver.py 3 > gpurun_out/s2f_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 6 -c 12 -o gpurun_out/s2f_full -f python tools/prof_driver.py 3 > gpurun_out/s2f_ncu_full.log 2>&1
python tools/prof_driver.py 3 4 4 1 > gpurun_out/s2f_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_analyze_sobel -s 1 -c 1 -o gpurun_out/s2f_sobel -f python tools/prof_driver.py 3 4 4 1 > gpurun_out/s2f_ncu_sobel.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2f_pytest.txt 2>&1; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2f_smoke.txt 2>&1; echo rc=$?
