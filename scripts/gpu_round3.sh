#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/*.log $O/status.txt
export PATH=/usr/local/cuda/bin:$PATH
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
timeout 200 $L/linkbench 1024 > $O/linkbench.json 2> $O/linkbench.err; echo "linkbench rc=$?" >> $O/status.txt
timeout 400 python scripts/pack_sweep.py > $O/pack_sweep.log 2>&1; echo "sweep rc=$?" >> $O/status.txt
# limiter diagnosis on the cudart/cuBLAS app
( export CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/lim.cache VGPU_PRINT_STATS=1 LIBCUDA_LOG_LEVEL=3 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 $L/gemm_loop 4096 4 ) > $O/gemm_lim.json 2> $O/gemm_lim.err; echo "gemm rc=$?" >> $O/status.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
( export CUDA_OVERSUBSCRIBE=true CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/sb.cache VGPU_PRINT_STATS=1 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 300 $L/swap_bench --cubin $CUBIN --buffers 384 --mib 64 --steps 768 --warmup 64 --profile 1 ) > $O/swap_small.json 2> $O/swap_small.err; echo "swap rc=$?" >> $O/status.txt
timeout 600 python bench.py --steps 24 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/status.txt
cat $O/status.txt; cat $O/linkbench.json; tail -3 $O/pack_sweep.log; cat $O/gemm_lim.json; grep "stats\|ERROR\|Warn" $O/gemm_lim.err | tail; tail -25 $O/pytest_gpu.log; cat $O/swap_small.json; tail -3 $O/swap_small.err; cat $O/bench.json; tail -5 $O/bench.err
