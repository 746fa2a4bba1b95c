#!/bin/bash
# r02 call 1: bring-up of the INT8 kernel + first hardware run of the --workload legs
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_r02a.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_r02a.txt
timeout 1300 python tools/ozaki_probe.py > $OUT/ozaki_probe_r02a.log 2>&1
echo "ozaki_probe rc=$?" >> $OUT/ozaki_probe_r02a.log
for w in ghz3 qft10 qft26 rqc6x6; do
  timeout 400 python bench.py --workload $w > $OUT/bench_r02a_$w.json 2> $OUT/bench_r02a_$w.err
  echo "$w rc=$?" >> $OUT/bench_r02a_rc.txt
done
timeout 400 python bench.py --workload qft26 --dtype c64 > $OUT/bench_r02a_qft26_c64.json 2> $OUT/bench_r02a_qft26_c64.err
timeout 400 python bench.py --workload rqc6x6 --dtype c64 > $OUT/bench_r02a_rqc6x6_c64.json 2> $OUT/bench_r02a_rqc6x6_c64.err
tail -40 $OUT/ozaki_probe_r02a.log
cat $OUT/bench_r02a_rc.txt
