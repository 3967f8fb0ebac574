#!/bin/bash
# ncu --set full captures of the second-generation kernel-pooling kernel and of the flat-IP kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_pool_ts -s 3 -c 1 -f -o gpurun_out/prof_kp_ts \
    python bench.py --workload tk --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_kp_ts.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_ip_tc -s 1 -c 1 -f -o gpurun_out/prof_flatip2 \
    python tests/tools/gpu_debug_flat_ip.py timing1 > gpurun_out/ncu_flatip2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_tk.csv \
    python bench.py --workload tk --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_tk.log 2>&1
ls -la gpurun_out/*.ncu-rep
