set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/s2_bench_n2b.json 2> gpurun_out/s2_bench_n2b.err; echo rc=$?
