#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 900 python -m pytest tests/test_gpu_hook.py tests/test_gpu_limiter.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; tail -30 $O/pytest_gpu.log
