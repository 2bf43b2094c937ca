#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
printf 'A 0 1048576\nI\nA 1 4194304\nF 0\nX 0x1234\nF 1\nT\nL 1 1 1\n' > /tmp/t8.txt
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ref_real.cache LIBCUDA_LOG_LEVEL=2 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 60 oracle/_ref/trace_replay /tmp/t8.txt ) > $O/ref_real.out 2> $O/ref_real.err; echo "ref trace rc=$?" >> $O/status.txt
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/new_real.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 oracle/_ref/trace_replay /tmp/t8.txt ) > $O/new_real.out 2>/dev/null
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/ref_ib.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 200 $L/intercept_bench $CUBIN 3000 100000 ) > $O/intercept_ref.json 2>/dev/null; echo "ref intercept rc=$?" >> $O/status.txt
timeout 60 $L/intercept_bench $CUBIN 3000 100000 > $O/intercept_bare.json 2>/dev/null
( export CUDA_DEVICE_MEMORY_LIMIT_0=8192m CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/new_ib.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 $L/intercept_bench $CUBIN 3000 100000 ) > $O/intercept_new.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/status.txt
timeout 400 python -m pytest tests/test_gpu_hook.py -m gpu -q --timeout 200 -p no:cacheprovider -k "reference_binary or hard_cap" > $O/pytest_ref.log 2>&1; echo "pytest rc=$?" >> $O/status.txt
cat $O/status.txt; echo REF; cat $O/ref_real.out; tail -3 $O/ref_real.err; echo NEW; cat $O/new_real.out; cat $O/intercept_bare.json $O/intercept_new.json $O/intercept_ref.json; cat $O/bench_ref.json; tail -3 $O/bench_ref.err; tail -5 $O/pytest_ref.log
