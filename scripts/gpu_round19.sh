#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
mkdir -p /tmp/vgpulock
timeout 400 python -m pytest tests/test_gpu_zz_pytorch.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_torch.log 2>&1; echo "torch rc=$?" >> $O/status.txt
timeout 300 python -m pytest tests/test_gpu_hook.py tests/test_gpu_limiter.py -m gpu -q --timeout 200 -p no:cacheprovider -k "launch_loop or replayed_cuda_graph or cublas or limiter" > $O/pytest_lim.log 2>&1; echo "lim rc=$?" >> $O/status.txt
cat $O/status.txt; tail -25 $O/pytest_torch.log | cut -c1-700; tail -12 $O/pytest_lim.log | cut -c1-500
