set -x
for so in var_ex128 var_ex384 var_ex512 var_ex128 var_ex384 var_ex512; do
PXZ_LIB=$PWD/pixlzr-rust_b200/$so.so python bench.py --steps 20 --warmup 5 --skip-extras --e2e-steps 1 --no-pcie-probe 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$so', round(d['value']), d['ms_per_step'], d['roofline']['single_stream_value_MPps'], {k:v['us'] for k,v in d['kernels'].items()})
" >> gpurun_out/s2_var12.txt
done
