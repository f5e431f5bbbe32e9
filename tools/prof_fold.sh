#!/bin/bash
# ncu captures of the folded specialised kernels (plain run first, then the captures)
mkdir -p gpurun_out
python tools/prof_model.py cartpole 65536 --lin > gpurun_out/r03q_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_linearize -c 1 -f -o gpurun_out/r03q_cart_lin python tools/prof_model.py cartpole 65536 --lin > gpurun_out/r03q_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -c 1 -f -o gpurun_out/r03q_drone_step python tools/prof_model.py drone 262144 > gpurun_out/r03q_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -c 1 -f -o gpurun_out/r03q_cart_step python tools/prof_model.py cartpole 65536 > gpurun_out/r03q_ncu3.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r03q_bench_plain.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r03q_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r03q_ncu4.log 2>&1
ls -la gpurun_out/r03q_*
