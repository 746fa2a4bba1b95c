#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; T=r02f
timeout 300 python tools/slice_breakdown.py > $OUT/slice_breakdown_${T}_c128.txt 2>&1
timeout 300 python tools/slice_breakdown.py c64 > $OUT/slice_breakdown_${T}_c64.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_${T}_n1.json 2> $OUT/bench_${T}_n1.err
timeout 600 python bench.py --dtype c64 --no-cpu-baseline > $OUT/bench_${T}_n1_c64.json 2> $OUT/bench_${T}_n1_c64.err
head -40 $OUT/slice_breakdown_${T}_c128.txt
head -30 $OUT/slice_breakdown_${T}_c64.txt
for f in $OUT/bench_${T}_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value')); print({k:(round(v.get('busy_ms_per_slice'),3),v.get('launches_per_slice')) for k,v in d.get('kernels',{}).items() if isinstance(v,dict) and 'busy_ms_per_slice' in v}); print(d.get('hoisted'))
"; done
