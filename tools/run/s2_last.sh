set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_last_pytest.txt 2>&1; echo rc=$?
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_last_smoke.txt 2>&1; echo rc=$?
python bench.py --steps 20 --warmup 5 > gpurun_out/s2_last_bench.json 2> gpurun_out/s2_last_bench.err; echo rc=$?
