mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-200
