mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
HPB_BENCH_STEP_TIMES=1 timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-clocks > gpurun_out/ab3.json 2> gpurun_out/ab3.err
grep "wall ms" gpurun_out/ab3.err
