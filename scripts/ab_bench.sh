run() { # label libenv workload
  MMB200_LIB=$2 timeout 200 python bench.py --workload $3 --no-secondary --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], '%.4g' % d['value'], 'frac %.4f' % d['roofline']['frac'], 'skip %.4g' % d.get('skip_padding',{}).get('value',0))"
}
timeout 200 python -m pytest tests/test_maxsim_gpu.py -m gpu -x -q 2>&1 | tail -3
PREV=/root/repo/scripts/ab_prev_libmatchmaker_b200.so
for i in 1 2; do run prev $PREV colbert; run cur "" colbert; done
for w in tk knrm; do run prev $PREV $w; run cur "" $w; run prev $PREV $w; run cur "" $w; done
