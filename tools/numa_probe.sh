mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)|thread"; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done; cat /sys/devices/system/node/node*/cpulist; free -g | head -2; } > gpurun_out/r2_numa_topo.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --skip-extras > gpurun_out/r2_numa_on.json 2> gpurun_out/r2_numa_on.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 4 --warmup 3 --skip-extras --no-numa-bind > gpurun_out/r2_numa_off.json 2> gpurun_out/r2_numa_off.err
tail -c 600 gpurun_out/r2_numa_on.json
