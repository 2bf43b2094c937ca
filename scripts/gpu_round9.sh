#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt $O/ab.log
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
timeout 600 python -m pytest tests/test_gpu_swap.py tests/test_gpu_hook.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not sm_limit" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
run_swap() { name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$$.cache VGPU_PRINT_STATS=1 LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 120 $L/swap_bench --cubin $CUBIN --buffers 192 --mib 64 --steps 512 --warmup 64 --profile 0 --verify 1 ) > $O/ab_tmp.json 2> $O/ab_tmp.err
  python3 -c "
import json
d=json.load(open('$O/ab_tmp.json')); print('$name', 'GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'], 'mism', d['mismatches'], 'creates', d['phys_creates'])" >> $O/ab.log; grep "slabs=" $O/ab_tmp.err | sed 's/.*ringwait=[0-9.]* //' >> $O/ab.log; rm -f /tmp/sb_$$.cache
}
for rep in 1 2 3 4; do
  run_swap async
  run_swap sync VGPU_SWAP_ASYNC_UNMAP=0
done
cat $O/status.txt; tail -8 $O/pytest_gpu.log; cat $O/ab.log
