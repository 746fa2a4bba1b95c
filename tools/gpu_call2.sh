#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python tools/ozaki_probe.py > $OUT/ozaki_probe_r02b.log 2>&1
echo "ozaki_probe rc=$?" >> $OUT/ozaki_probe_r02b.log
tail -150 $OUT/ozaki_probe_r02b.log
