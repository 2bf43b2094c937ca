#!/bin/bash
# Round-end pass: full GPU test suite, the bench (both arms), ncu launch list of the bench's timed region, one full
# capture of the dominant kernel, victim-scan scaling. Outputs under gpurun_out/, summarised into profiles/ afterwards.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
export PATH=/usr/local/cuda/bin:$PATH
mkdir -p /tmp/vgpulock
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
timeout 300 python bench.py --impl reference --steps 4 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/status.txt
timeout 400 ncu --target-processes application-only --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --skip-cpu-baseline > $O/ncu_bench.log 2>&1; echo "ncu bench rc=$?" >> $O/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python scripts/ncu_target.py > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/status.txt
TOUCHES=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vgpu_pack_tma -s 40 -c 3 -o $O/prof_pack -f python scripts/ncu_target.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/status.txt
timeout 200 python scripts/scan_scaling.py > $O/scan_scaling.log 2>&1; echo "scan rc=$?" >> $O/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_victim --csv --log-file $O/scan_launches_1m.csv python scripts/scan_scaling.py 1048576 > $O/scan_ncu.log 2>&1; echo "ncu scan rc=$?" >> $O/status.txt
cat $O/status.txt; tail -12 $O/pytest_gpu.log; cat $O/bench.json; tail -3 $O/bench.err; cat $O/bench_ref.json; grep -c vgpu_ $O/launches_bench.csv; tail -3 $O/ncu_bench.log; tail -7 $O/scan_scaling.log
