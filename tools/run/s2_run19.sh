set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest19.txt 2>&1; echo rc=$?
bash tools/variant_times.sh > gpurun_out/s2_var19.txt 2>&1
for so in libpixlzr_b200 var_nopipe libpixlzr_b200 var_nopipe; do
PXZ_LIB=$PWD/pixlzr-rust_b200/$so.so python bench.py --steps 20 --warmup 5 --e2e-steps 1 --no-pcie-probe --c5-images 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$so', round(d['value']), d['ms_per_step'], d['roofline']['single_stream_value_MPps'], {k:v['us'] for k,v in d['kernels'].items()}, 'c4norm', d['sharded']['normalise_global']['MPps'])
" >> gpurun_out/s2_var19.txt
done
