set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest13.txt 2>&1; echo rc=$?
for st in 4 6 8; do
python bench.py --steps 20 --warmup 5 --skip-extras --e2e-steps 1 --no-pcie-probe --streams $st --batch 24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams $st batch 24', round(d['value']), d['ms_per_step'])
" >> gpurun_out/s2_var13.txt
done
python bench.py --steps 20 --warmup 5 --skip-extras --e2e-steps 1 --no-pcie-probe --streams 3 --batch 18 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('streams 3 batch 18', round(d['value']), d['ms_per_step'])
" >> gpurun_out/s2_var13.txt
