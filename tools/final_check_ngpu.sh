#!/bin/bash
# bench.py at N GPUs, launched the way the driver does:  tools/final_check_ngpu.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N > gpurun_out/bench_final_${N}gpu.json 2> gpurun_out/bench_final_${N}.err
tail -c 200 gpurun_out/bench_final_${N}gpu.json; echo
