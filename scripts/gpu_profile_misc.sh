#!/bin/bash
# ncu captures of the flat-IP and kernel-pool kernels + backward timing
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_ip_tc -s 1 -c 1 -f -o gpurun_out/prof_flatip \
    python tests/tools/gpu_debug_flat_ip.py timing1 > gpurun_out/ncu_flatip.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_pool_tc -s 3 -c 1 -f -o gpurun_out/prof_kp2 \
    python bench.py --workload tk --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_kp2.log 2>&1
timeout 300 python tests/tools/gpu_debug_kp.py bwd
