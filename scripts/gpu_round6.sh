#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.log $O/status.txt $O/engine_probe.log
export PATH=/usr/local/cuda/bin:$PATH
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
timeout 120 $L/vmm_dma_probe > $O/vmm_dma_probe.json 2> $O/vmm_dma_probe.err; echo "probe rc=$?" >> $O/status.txt
timeout 600 python -m pytest tests/test_gpu_swap.py tests/test_gpu_kernels.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
run_swap() { name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$name.cache LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 1 ) > $O/swap_$name.json 2> $O/swap_$name.err; echo "swap $name rc=$?" >> $O/status.txt
}
run_swap default
run_swap ring8 VGPU_SWAP_RING=8
run_swap chunk16 VGPU_SWAP_CHUNK_MB=16 VGPU_SWAP_RING=8
for vnt in plain plain_prof; do timeout 200 python scripts/engine_probe.py $vnt >> $O/engine_probe.log 2>&1; done
timeout 600 python bench.py --steps 24 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; cat $O/vmm_dma_probe.json; tail -6 $O/pytest_gpu.log; for n in default ring8 chunk16; do echo $n; python3 -c "
import json
d=json.load(open('$O/swap_$n.json')); print('GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1), d['host_ms'], 'pack ev/span ms', d['pack_ms'], d['pack_span_ms'], 'unpack', d['unpack_ms'], d['unpack_span_ms'], 'launches', d['pack_launches'], 'mism', d['mismatches'])"; done; grep -v "^numa" $O/engine_probe.log; cat $O/bench.json; tail -3 $O/bench.err
