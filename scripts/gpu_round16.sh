#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 200 python scripts/scan_scaling.py > $O/scan_scaling.log 2>&1; echo "scan rc=$?" >> $O/status.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_victim --csv --log-file $O/scan_launches_small.csv python scripts/scan_scaling.py 1152 8192 > $O/scan_ncu.log 2>&1; echo "ncu scan rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 48 --warmup 4 --skip-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; tail -8 $O/pytest_gpu.log; tail -7 $O/scan_scaling.log; grep -v "^==" $O/scan_launches_small.csv | tail -4 | cut -c1-200; cat $O/bench.json | cut -c1-1200
