#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 600 python -m pytest tests/test_gpu_hook.py -m gpu -q --timeout 300 -p no:cacheprovider -k "async_pool or replayed_cuda_graph or captured_kernels or driver_api_launch_loop" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; tail -40 $O/pytest_gpu.log
