set -x
for amp in 16 4; do
ncu --set full --clock-control none --import-source on -k regex:'k_shrink_tma|k_expand_warp' -s 4 -c 2 -o gpurun_out/s2b_lvl_amp$amp -f python tools/prof_level.py $amp 3 > gpurun_out/s2b_ncu_amp$amp.log 2>&1
done
