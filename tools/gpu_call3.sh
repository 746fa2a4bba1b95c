#!/bin/bash
# r02 call 3: re-confirm the default path, headline with/without the INT8 kernels, ncu of the INT8 kernels
OUT=gpurun_out; mkdir -p $OUT; T=r02c
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$T.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_$T.txt
PQ_TEST_OZAKI=1 timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_$T.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$T.log
timeout 600 python bench.py > $OUT/bench_${T}_n1.json 2> $OUT/bench_${T}_n1.err
timeout 600 python bench.py --ozaki 6 --no-cpu-baseline > $OUT/bench_${T}_n1_ozaki6.json 2> $OUT/bench_${T}_n1_ozaki6.err
timeout 600 python bench.py --dtype c64 --no-cpu-baseline > $OUT/bench_${T}_n1_c64.json 2> $OUT/bench_${T}_n1_c64.err
timeout 600 python bench.py --dtype c64 --ozaki 4 --no-cpu-baseline > $OUT/bench_${T}_n1_c64_ozaki4.json 2> $OUT/bench_${T}_n1_c64_ozaki4.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_ozaki -s 2 -c 1 \
  -o $OUT/prof_${T}_ozaki6 -f python tools/ozaki_one.py c128 6 > $OUT/ncu_${T}_ozaki6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_ozaki -s 2 -c 1 \
  -o $OUT/prof_${T}_ozaki4 -f python tools/ozaki_one.py c64 4 > $OUT/ncu_${T}_ozaki4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_skinny -s 2 -c 1 \
  -o $OUT/prof_${T}_skinny -f python tools/ozaki_one.py c128 0 > $OUT/ncu_${T}_skinny.log 2>&1
tail -5 $OUT/pytest_$T.log
for f in $OUT/bench_${T}_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('roofline'))
except Exception as e: print('ERR', e)
"; done
tail -3 $OUT/*.err
ls -la $OUT | tail -30
