#!/bin/bash
# usage (on the GPU box): bash scripts/run_scaling_suite.sh N [extra]   -> gpurun_out/scale_n<N>_*.json
# the driver's launch line for N ranks, weak (1 detection per GPU) and strong (one frame's 576 rows over N ranks)
N=$1; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n${N}_weak.json 2> gpurun_out/scale_n${N}_weak.err
timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --scaling strong > gpurun_out/scale_n${N}_strong.json 2> gpurun_out/scale_n${N}_strong.err
if [ "$2" = "configs" ]; then
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --config tless240 > gpurun_out/scale_n${N}_tless240.json 2> gpurun_out/scale_n${N}_tless240.err
  timeout 900 $TR --master-port 29514 bench.py --gpus $N --config gso1000 > gpurun_out/scale_n${N}_gso1000.json 2> gpurun_out/scale_n${N}_gso1000.err
fi
for f in gpurun_out/scale_n${N}_*.json; do echo $f; cut -c1-200 $f; done
for f in gpurun_out/scale_n${N}_*.err; do tail -n 3 $f; done
