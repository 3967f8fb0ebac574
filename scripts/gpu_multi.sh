#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/tools/gpu_multi_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -6
for w in colbert bert_dot; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_${w}_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$w n=$N", "value=%.4g" % d["value"], "ms_per_step=%.3f" % d["ms_per_step"], "kern_ms=%.3f" % d["roofline"]["kernel_ms"], "e2e=%.4g" % d["e2e"]["value"])
except Exception as e:
    print("$w n=$N ERR", e); print(open("gpurun_out/bench_${w}_n$N.err").read()[-1500:])
PY
done
