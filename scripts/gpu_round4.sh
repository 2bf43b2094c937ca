#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.log $O/status.txt
export PATH=/usr/local/cuda/bin:$PATH
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
# how long is one cuBLAS SGEMM 4096^3 really? (ncu, no hook)
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 3 -c 5 --csv --log-file $O/gemm_ncu.csv $L/gemm_loop 4096 0.05 > $O/gemm_ncu.log 2>&1; echo "gemm ncu rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
run_swap() { # name, extra env...
  name=$1; shift
  ( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_$name.cache LD_PRELOAD=$PWD/$L/libvgpu.so "$@"; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 1 ) > $O/swap_$name.json 2> $O/swap_$name.err; echo "swap $name rc=$?" >> $O/status.txt
}
run_swap default
run_swap nonuma VGPU_SWAP_NO_NUMA=1
run_swap chunk64 VGPU_SWAP_CHUNK_MB=64
run_swap ring8 VGPU_SWAP_RING=8
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb_np.cache LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 0 ) > $O/swap_noprofile.json 2> $O/swap_noprofile.err
timeout 600 python bench.py --steps 24 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python scripts/ncu_target.py > $O/ncu_list.log 2>&1; echo "ncu list rc=$?" >> $O/status.txt
TOUCHES=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vgpu_pack_tma -s 40 -c 3 -o $O/prof_pack python scripts/ncu_target.py > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/status.txt
cat $O/status.txt; grep -v "^==" $O/gemm_ncu.csv | tail -6; tail -25 $O/pytest_gpu.log; for n in default nonuma chunk64 ring8 noprofile; do echo $n; python3 -c "
import json,sys
d=json.load(open('$O/swap_$n.json')); print({k:d[k] for k in ('event_ms','pack_ms','unpack_ms','pack_launches','scans','scan_cache_hits','phys_reuses','mismatches')}, 'GB/s', round((d['page_in_bytes']+d['page_out_bytes'])/d['event_ms']/1e6,1))"; done; cat $O/bench.json; tail -5 $O/bench.err
