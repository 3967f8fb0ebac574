#!/bin/bash
# Same-box A/B of two builds of the library (bench.py value lines): MMB200_LIB selects the build.
#   scripts/ab_bench.sh <other .so> <workload> [<workload> ...]
PREV=$1; shift
run() { # label lib workload
  MMB200_LIB=$2 timeout 300 python bench.py --workload $3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], '%.4g' % d['value'], 'frac %.4f' % d['roofline']['frac'], 'ms %.4f' % d['ms_per_step'])"
}
for w in "$@"; do for i in 1 2; do run other $PREV $w; run tree "" $w; done; done
