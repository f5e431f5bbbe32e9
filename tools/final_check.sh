#!/bin/bash
# Round-end check on one GPU: GPU suite, smoke(), both bench arms (outputs under gpurun_out/)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final_1gpu.json; echo
python bench.py --impl reference > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final_reference.json
