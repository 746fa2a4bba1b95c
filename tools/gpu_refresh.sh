#!/bin/bash
# One gpurun call that refreshes every measured artefact (run from the repo root on the GPU box):
#
#   gpurun --timeout 2400 -- 'bash tools/gpu_refresh.sh r02z'
#
# then, back in the container:
#   python tools/ncu_summary.py r02z gpurun_out/launches_r02z.csv gpurun_out/prof_r02z_ozaki_t.ncu-rep
#   cp gpurun_out/bench_r02z_*.json profiles/
#
# Every leg runs under its own `timeout`, so a hung kernel cannot hold the box until gpurun's
# limit.  Numbers printed by the runs under ncu are never bench values.
TAG=${1:-refresh}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu_$TAG.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/gpu_$TAG.txt

# 1. parity first
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_$TAG.log

# 2. the headline (BASELINE config 5), both element types, the A/B without the INT8 kernel, and
#    the reference arm beside it
timeout 600 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
timeout 600 python bench.py --dtype c64 > $OUT/bench_${TAG}_n1_c64.json 2> $OUT/bench_${TAG}_n1_c64.err
timeout 600 python bench.py --no-int8 --no-cpu-baseline > $OUT/bench_${TAG}_n1_noint8.json 2> $OUT/bench_${TAG}_n1_noint8.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_ref.json 2> $OUT/bench_${TAG}_ref.err

# 3. BASELINE configs 1-4 through the same JSON contract
for w in ghz3 qft10 qft26 rqc6x6; do
  timeout 600 python bench.py --workload $w > $OUT/bench_${TAG}_$w.json 2> $OUT/bench_${TAG}_$w.err
done
timeout 600 python bench.py --workload qft26 --dtype c64 > $OUT/bench_${TAG}_qft26_c64.json 2> $OUT/bench_${TAG}_qft26_c64.err
timeout 600 python bench.py --workload rqc6x6 --dtype c64 > $OUT/bench_${TAG}_rqc6x6_c64.json 2> $OUT/bench_${TAG}_rqc6x6_c64.err

# 4. ncu: launch list of the bench command, then full counters of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --lanes 1 --no-cpu-baseline --no-profile > $OUT/ncu_launch_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_t -s 2 -c 1 \
  -o $OUT/prof_${TAG}_ozaki_t -f python tools/ozaki_one.py c128 0 > $OUT/ncu_full_${TAG}_ozaki_t.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_t -s 2 -c 1 \
  -o $OUT/prof_${TAG}_ozaki_t_c64 -f python tools/ozaki_one.py c64 0 > $OUT/ncu_full_${TAG}_ozaki_t_c64.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_zgemm_skinny -s 2 -c 1 \
  -o $OUT/prof_${TAG}_skinny -f python tools/ozaki_one.py c128 -1 > $OUT/ncu_full_${TAG}_skinny.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cgemm_tcgen05 -s 1 -c 1 \
  -o $OUT/prof_${TAG}_cgemm -f python tools/cgemm_probe.py > $OUT/ncu_full_${TAG}_cgemm.log 2>&1

# 5. per-shape probes: the INT8 kernel against the other path, the MMA / TMEM / FP64 rate probes
timeout 600 python tools/ozaki_t_probe.py > $OUT/ozaki_t_probe_$TAG.log 2>&1
cp $OUT/ozaki_t_probe.json $OUT/ozaki_t_probe_$TAG.json
timeout 120 python tools/ozaki_t_rate.py 0 1 2 4 8 16 32 64 256 512 768 1536 2048 2304 2560 2816 > $OUT/ozaki_t_rate_$TAG.log 2>&1
timeout 120 python tools/ozaki_t_ldtm.py > $OUT/ozaki_t_ldtm_$TAG.log 2>&1
timeout 300 python tools/slice_breakdown.py > $OUT/slice_breakdown_${TAG}_c128.txt 2>&1
timeout 300 python tools/slice_breakdown.py c64 > $OUT/slice_breakdown_${TAG}_c64.txt 2>&1
timeout 300 python tools/cgemm_probe.py > $OUT/cgemm_probe_$TAG.log 2>&1
# thin-N DMMA kernel against the tiled one: kernel durations from an ncu launch list
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/thin_launches_$TAG.csv \
  -k regex:k_zgemm python tools/thin_probe.py > $OUT/thin_probe_$TAG.log 2>&1
ls -la $OUT | tail -40
