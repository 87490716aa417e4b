python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 1 --no-pcie-probe --c4-side 16384 --c5-images 256 > gpurun_out/s2_n2q.json 2> gpurun_out/s2_n2q.err; echo rc=$?
tail -2 gpurun_out/s2_n2q.err
