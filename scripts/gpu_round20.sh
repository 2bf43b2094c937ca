#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
mkdir -p /tmp/vgpulock
timeout 100 python -m pytest tests/test_gpu_zz_pytorch.py -m gpu -q --timeout 90 -p no:cacheprovider -k oversubscribes > $O/pytest_torch_swap.log 2>&1; echo "rc=$?"
tail -15 $O/pytest_torch_swap.log | cut -c1-600
