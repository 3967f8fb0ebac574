#!/bin/bash
# Run on the GPU box: full GPU test suite, smoke, and one bench line per workload.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
for w in colbert tk knrm tkl bert_dot; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  timeout 600 python bench.py --impl reference --workload $w --steps 2 --warmup 1 > gpurun_out/bench_ref_$w.json 2>> gpurun_out/bench_$w.err
done
cat gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
for w in colbert tk knrm tkl bert_dot; do python - <<PY
import json
for f in ("gpurun_out/bench_$w.json", "gpurun_out/bench_ref_$w.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "value=%.4g" % d["value"], "e2e=%.4g" % d["e2e"]["value"], "frac=%s" % r.get("frac"), "kern_ms=%s" % r.get("kernel_ms"), "cpu=%s" % (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json", ".err").replace("_ref", "")).read()[-800:])
PY
done
