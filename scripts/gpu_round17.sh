#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 500 python -m pytest tests/test_gpu_zz_pytorch.py -m gpu -q --timeout 240 -p no:cacheprovider -x > $O/pytest_torch.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; tail -40 $O/pytest_torch.log | cut -c1-600
