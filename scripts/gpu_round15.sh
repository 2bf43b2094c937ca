#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -p no:cacheprovider > $O/pytest_kernels.log 2>&1; echo "kernels rc=$?" >> $O/status.txt
timeout 200 python scripts/scan_scaling.py > $O/scan_scaling.log 2>&1; echo "scan rc=$?" >> $O/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:vgpu_victim --csv --log-file $O/scan_launches_1m.csv python scripts/scan_scaling.py 1048576 > $O/scan_ncu.log 2>&1; echo "ncu rc=$?" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_hook.py -m gpu -q --timeout 200 -p no:cacheprovider -k "node_monitor_sees" > $O/pytest_hook.log 2>&1; echo "hook rc=$?" >> $O/status.txt
cat $O/status.txt; tail -5 $O/pytest_kernels.log; cat $O/scan_scaling.log | tail -8; tail -5 $O/pytest_hook.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/scan_launches_1m.csv")) if len(r) > 5 and r[0].isdigit()]
# last complete scan = last 7 launches (init, 4 hist, count, emit)
agg = collections.OrderedDict()
for r in rows[-7:]:
    print(r[4][:40], r[-1], r[-2])
PY
