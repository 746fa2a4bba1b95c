#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; T=r02e
PQ_TEST_OZAKI=1 timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$T.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$T.log
timeout 600 python bench.py > $OUT/bench_${T}_n1.json 2> $OUT/bench_${T}_n1.err
timeout 600 python bench.py --dtype c64 --no-cpu-baseline > $OUT/bench_${T}_n1_c64.json 2> $OUT/bench_${T}_n1_c64.err
tail -8 $OUT/pytest_$T.log
for f in $OUT/bench_${T}_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value')); print(d.get('roofline')); print({k:(v.get('busy_ms_per_slice'),v.get('launches_per_slice')) for k,v in d.get('kernels',{}).items() if isinstance(v,dict) and 'busy_ms_per_slice' in v}); print(d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
"; done
tail -n 3 $OUT/bench_${T}_n1.err $OUT/bench_${T}_n1_c64.err
