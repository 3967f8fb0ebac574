#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload colbert --steps 20 --warmup 5 > gpurun_out/bench_colbert.json 2> gpurun_out/bench_colbert.err
timeout 600 python bench.py --workload tk --steps 10 --warmup 3 > gpurun_out/bench_tk.json 2> gpurun_out/bench_tk.err
timeout 600 python bench.py --workload knrm --steps 10 --warmup 3 > gpurun_out/bench_knrm.json 2> gpurun_out/bench_knrm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kernel_pool_tc -s 3 -c 1 -f -o gpurun_out/prof_kp \
    python bench.py --workload tk --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_kp.log 2>&1
cat gpurun_out/pytest_gpu.log
for w in colbert tk knrm; do python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("$w", "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "frac=%.3f" % r["frac"], "kern_ms=%.3f" % r["kernel_ms"], "cpu=%s" % (d.get("cpu_baseline") or {}).get("value"), "clocks=%s" % d.get("clocks"), "skip=%s" % (d.get("skip_padding") or {}).get("value"))
PY
done
