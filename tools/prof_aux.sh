#!/bin/bash
# Throughput of b2_dlqr and one ncu capture of the DARE kernels and of the small per-step kernels of the
# time-varying LQR / random-control / recorder paths (run under gpurun; outputs in gpurun_out/).
set -u
mkdir -p gpurun_out
python tools/dlqr_bench.py > gpurun_out/dlqr_bench.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_dare -c 3 -f -o gpurun_out/dare \
    python tools/dlqr_bench.py > gpurun_out/ncu_dare.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k_lqr_control_env|k_random_controls|k_record_rows|k_dare|k_commit_state|k_cost' -c 60 --csv \
    --log-file gpurun_out/aux_kernels.csv python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -q -m gpu -k "tv_lqr or random_controller or recorder or dare" \
    > gpurun_out/ncu_aux.log 2>&1
tail -3 gpurun_out/ncu_aux.log
cat gpurun_out/dlqr_bench.txt
