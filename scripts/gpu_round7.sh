#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
timeout 200 $L/vmm_dma_probe > $O/vmm_dma_probe.json 2> $O/vmm_dma_probe.err; echo "probe rc=$?" >> $O/status.txt
timeout 100 $L/linkbench 1024 > $O/linkbench.json 2>/dev/null
run_swap() { name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$name.cache LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 512 --warmup 64 --profile 0 ) > $O/swap_$name.json 2> $O/swap_$name.err; echo "swap $name rc=$?" >> $O/status.txt
}
run_swap a
run_swap b
run_swap chunk64 VGPU_SWAP_CHUNK_MB=64
run_swap look32 VGPU_SWAP_SCAN_LOOKAHEAD=32
run_swap arena64 VGPU_SWAP_ARENA_GB=64
cat $O/status.txt; cat $O/vmm_dma_probe.json; cat $O/linkbench.json; for n in a b chunk64 look32 arena64; do echo $n; python3 -c "
import json
d=json.load(open('$O/swap_$n.json')); print('GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'], 'scans', d['scans'])"; done
