set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/s2_bench_n8.json 2> gpurun_out/s2_bench_n8.err; echo rc=$?
tail -3 gpurun_out/s2_bench_n8.err
