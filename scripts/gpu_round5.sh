#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.log $O/status.txt
export PATH=/usr/local/cuda/bin:$PATH
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
for vnt in plain bind bind_smi bind_prof bind_stream; do timeout 200 python scripts/engine_probe.py $vnt >> $O/engine_probe.log 2>&1; done
run_swap() { name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$name.cache VGPU_PRINT_STATS=1 LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 0 ) > $O/swap_$name.json 2> $O/swap_$name.err; echo "swap $name rc=$?" >> $O/status.txt
}
run_swap default
run_swap default2
( taskset -c 32-63 true ) 2>/dev/null && { export -f run_swap 2>/dev/null; }
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_far.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 taskset -c 32-63 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 0 ) > $O/swap_farcpu.json 2> $O/swap_farcpu.err
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_near.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 taskset -c 0-31 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 0 ) > $O/swap_nearcpu.json 2> $O/swap_nearcpu.err
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_farnn.cache VGPU_SWAP_NO_NUMA=1 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 taskset -c 32-63 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 0 ) > $O/swap_farcpu_nonuma.json 2> $O/swap_farcpu_nonuma.err
timeout 600 python -m pytest tests/test_gpu_hook.py tests/test_gpu_limiter.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; cat $O/engine_probe.log | grep -v "^numa"; for n in default default2 farcpu nearcpu farcpu_nonuma; do echo $n; python3 -c "
import json
d=json.load(open('$O/swap_$n.json')); print('GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'], 'enq', d['enqueue_ms'])"; done; grep stats $O/swap_default.err; tail -12 $O/pytest_gpu.log
