set -x
bash tools/variant_times.sh > gpurun_out/s2_var18.txt 2>&1
for so in libpixlzr_b200 var_noredux libpixlzr_b200 var_noredux; do
PXZ_LIB=$PWD/pixlzr-rust_b200/$so.so python bench.py --steps 20 --warmup 5 --skip-extras --e2e-steps 1 --no-pcie-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$so', round(d['value']), d['ms_per_step'], d['roofline']['single_stream_value_MPps'], {k:v['us'] for k,v in d['kernels'].items()})
" >> gpurun_out/s2_var18.txt
done
python -m pytest tests -m gpu -x -q -k "mad or analy or normalise or parity or 8k" > gpurun_out/s2_pytest18.txt 2>&1; echo rc=$?
