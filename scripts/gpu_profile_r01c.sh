#!/bin/bash
# ncu --set full capture of the final flat-IP kernel + launch list of the bert_dot bench step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_ip_tc -s 1 -c 1 -f -o gpurun_out/prof_flatip3 \
    python tests/tools/gpu_debug_flat_ip.py timing1 > gpurun_out/ncu_flatip3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_bert_dot.csv \
    python bench.py --workload bert_dot --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_bd.log 2>&1
ls -la gpurun_out/prof_flatip3.ncu-rep gpurun_out/launches_bert_dot.csv
