#!/bin/bash
# Round-end 8-GPU run of both bench arms, launched the way the driver does
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --impl reference > gpurun_out/bench_final_reference_8gpu.json 2> gpurun_out/bench_final8.err
tail -c 300 gpurun_out/bench_final_reference_8gpu.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 > gpurun_out/bench_final_8gpu.json 2>> gpurun_out/bench_final8.err
tail -c 300 gpurun_out/bench_final_8gpu.json; echo; tail -3 gpurun_out/bench_final8.err
