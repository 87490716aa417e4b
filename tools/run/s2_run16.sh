set -x
for so in libpixlzr_b200 var_notiny; do
PXZ_LIB=$PWD/pixlzr-rust_b200/$so.so ncu --set full --clock-control none -k regex:k_expand_warp -s 2 -c 1 -o gpurun_out/s2_mixed_$so -f python tools/prof_driver.py 3 > gpurun_out/s2_ncu16_$so.log 2>&1
done
