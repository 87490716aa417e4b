set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest14.txt 2>&1; echo rc=$?
timeout 600 python tests/tools/fuzz_parity.py 4000 77 > gpurun_out/s2_fuzz14.txt 2>&1; echo rc=$?
python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-pcie-probe > gpurun_out/s2_b14.json 2> gpurun_out/s2_b14.err; echo rc=$?
