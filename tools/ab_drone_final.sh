#!/bin/bash
# Final check of the drone k_step change: full GPU suite, ncu --set full capture the way profiles/ncu_drone_step_r02i.txt
# was made (tools/prof_model.py: plain b2_step, no parked state), default bench line.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:k_step -c 1 -f -o gpurun_out/r03m_drone_step python tools/prof_model.py drone 262144 > gpurun_out/r03m_ncu.log 2>&1
python bench.py > gpurun_out/bench_r03m_1gpu.json 2> gpurun_out/bench_r03m.err; tail -c 600 gpurun_out/bench_r03m_1gpu.json
