#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt $O/ab.log
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
timeout 200 $L/vmm_dma_probe > $O/vmm_dma_probe.json 2> $O/vmm_dma_probe.err; echo "probe rc=$?" >> $O/status.txt
run_swap() { name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$$.cache LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 120 $L/swap_bench --cubin $CUBIN --buffers 192 --mib 64 --steps 512 --warmup 64 --profile 0 --verify 0 ) > $O/ab_tmp.json 2>/dev/null
  python3 -c "
import json
d=json.load(open('$O/ab_tmp.json')); print('$name', 'GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'])" >> $O/ab.log; rm -f /tmp/sb_$$.cache
}
for rep in 1 2 3 4; do
  run_swap align4g
  run_swap align0 VGPU_SWAP_ARENA_ALIGN_GB=0
  run_swap arena16 VGPU_SWAP_ARENA_GB=16
done
cat $O/status.txt; cat $O/vmm_dma_probe.json; sort $O/ab.log
