#!/bin/bash
# One gpurun call that refreshes every measured artefact (run from the repo root on the GPU box):
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_refresh.sh r02a'
#
# then, back in the container:
#   python tools/ncu_summary.py r02a gpurun_out/launches_r02a.csv gpurun_out/prof_r02a.ncu-rep
#   cp gpurun_out/bench_r02a_*.json profiles/
#
# Every leg runs under its own `timeout`, so a hung kernel cannot hold the box until gpurun's
# limit.  Numbers printed by the runs under ncu are never bench values.
TAG=${1:-refresh}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1

# 1. parity first
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log

# 2. the headline (BASELINE config 5) and the reference arm beside it
timeout 600 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_ref.json 2> $OUT/bench_${TAG}_ref.err

# 3. BASELINE configs 1-4 through the same JSON contract
for w in ghz3 qft10 qft26 rqc6x6; do
  timeout 600 python bench.py --workload $w > $OUT/bench_${TAG}_$w.json 2> $OUT/bench_${TAG}_$w.err
done
timeout 600 python bench.py --workload qft26 --dtype c64 > $OUT/bench_${TAG}_qft26_c64.json 2> $OUT/bench_${TAG}_qft26_c64.err
timeout 600 python bench.py --workload rqc6x6 --dtype c64 > $OUT/bench_${TAG}_rqc6x6_c64.json 2> $OUT/bench_${TAG}_rqc6x6_c64.err

# 4. ncu: launch list of the bench command, then full counters of the dominant kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline --no-profile > $OUT/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_skinny -s 20 -c 3 \
  -o $OUT/prof_$TAG -f \
  python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline --no-profile > $OUT/ncu_full_$TAG.log 2>&1

# 5. per-shape probe of the sweep-step GEMMs (A/B of the kernel variants)
timeout 300 python tools/gemm_probe.py > $OUT/gemm_probe_$TAG.log 2>&1

# 6. bring-up of the experimental INT8 tensor-core ZGEMM (staged, each stage under a timeout)
timeout 1300 python tools/ozaki_probe.py > $OUT/ozaki_probe_$TAG.log 2>&1
OZ_RC=$?
echo "ozaki_probe rc=$OZ_RC" >> $OUT/ozaki_probe_$TAG.log
if [ $OZ_RC -eq 0 ]; then
  # 7. only after a green probe: gated parity tests, the headline with the INT8 kernel, its counters
  PQ_TEST_OZAKI=1 timeout 600 python -m pytest tests/test_gpu_ozaki.py -q > $OUT/pytest_ozaki_$TAG.log 2>&1
  for g in 6 7; do
    timeout 600 python bench.py --ozaki $g > $OUT/bench_${TAG}_n1_ozaki$g.json 2> $OUT/bench_${TAG}_n1_ozaki$g.err
  done
  timeout 600 python bench.py --workload rqc6x6 --ozaki 6 > $OUT/bench_${TAG}_rqc6x6_ozaki6.json 2> $OUT/bench_${TAG}_rqc6x6_ozaki6.err
  timeout 600 python bench.py --workload rqc6x6 --dtype c64 --ozaki 4 > $OUT/bench_${TAG}_rqc6x6_c64_ozaki4.json 2> $OUT/bench_${TAG}_rqc6x6_c64_ozaki4.err
  timeout 600 python bench.py --dtype c64 --ozaki 4 > $OUT/bench_${TAG}_n1_c64_ozaki4.json 2> $OUT/bench_${TAG}_n1_c64_ozaki4.err
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_ozaki -s 20 -c 3 \
    -o $OUT/prof_${TAG}_ozaki -f \
    python bench.py --ozaki 6 --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline --no-profile > $OUT/ncu_full_${TAG}_ozaki.log 2>&1
fi
ls -la $OUT | tail -30
