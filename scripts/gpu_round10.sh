#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
for i in 1 2; do
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_t$i.cache VGPU_PRINT_STATS=1 VGPU_SWAP_TRACE=120 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 120 $L/swap_bench --cubin $CUBIN --buffers 192 --mib 64 --steps 640 --warmup 64 --profile 0 --verify 0 ) > $O/trace$i.json 2> $O/trace$i.err; echo "trace$i rc=$?" >> $O/status.txt
done
cat $O/status.txt; for i in 1 2; do python3 -c "
import json
d=json.load(open('$O/trace$i.json')); print('GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'])"; done
