set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/s2_graph.json 2> gpurun_out/s2_graph.err; echo rc=$?
python bench.py --steps 20 --warmup 5 --issue direct --skip-extras > gpurun_out/s2_direct.json 2> gpurun_out/s2_direct.err; echo rc=$?
python tools/class_times.py > gpurun_out/s2_class_times.txt 2>&1
for amp in 0 2; do
ncu --set full --clock-control none --import-source on -k regex:'k_shrink_tma|k_expand_warp' -s 4 -c 2 -o gpurun_out/s2_lvl_amp$amp -f python tools/prof_level.py $amp 3 > gpurun_out/s2_ncu_amp$amp.log 2>&1
done
