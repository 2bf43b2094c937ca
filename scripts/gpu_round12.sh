#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O; rm -f $O/status.txt $O/cfg4.log
L=k8s-device-plugin_b200/lib; CUBIN=k8s-device-plugin_b200/build/vgpu_kernels.cubin
mkdir -p /tmp/vgpulock
smi() { nvidia-smi -i 0 --query-gpu=utilization.gpu --format=csv,noheader,nounits -lms 100 > $O/smi_$1.txt & echo $! > /tmp/smi.pid; }
endsmi() { kill $(cat /tmp/smi.pid); python3 -c "
v=[int(x) for x in open('$O/smi_$1.txt').read().split() if x.isdigit()]; b=v[len(v)//4:-max(1,len(v)//8)] or v; print('$1 smi_util_avg', round(sum(b)/max(len(b),1),1))" >> $O/cfg4.log; }
smi bare; timeout 60 $L/launch_loop $CUBIN 2048 6 >> $O/cfg4.log 2>&1; endsmi bare
smi new30; ( export CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/l_new.cache VGPU_PRINT_STATS=1 LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 $L/launch_loop $CUBIN 2048 6 ) >> $O/cfg4.log 2>&1; endsmi new30
smi ref30; ( export CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/l_ref.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 90 $L/launch_loop $CUBIN 2048 6 ) >> $O/cfg4.log 2>&1; endsmi ref30
smi new60; ( export CUDA_DEVICE_SM_LIMIT=60 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/l_new6.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 $L/launch_loop $CUBIN 2048 6 ) >> $O/cfg4.log 2>&1; endsmi new60
smi ref60; ( export CUDA_DEVICE_SM_LIMIT=60 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/l_ref6.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/oracle/_ref/dlsym_shim.so:$PWD/oracle/_ref/libvgpu.so; timeout 90 $L/launch_loop $CUBIN 2048 6 ) >> $O/cfg4.log 2>&1; endsmi ref60
smi new30small; ( export CUDA_DEVICE_SM_LIMIT=30 GPU_CORE_UTILIZATION_POLICY=force CUDA_DEVICE_MEMORY_SHARED_CACHE=/tmp/l_news.cache LIBCUDA_LOG_LEVEL=0 LD_PRELOAD=$PWD/$L/libvgpu.so; timeout 60 $L/launch_loop $CUBIN 16 6 ) >> $O/cfg4.log 2>&1; endsmi new30small
timeout 200 python scripts/scan_scaling.py > $O/scan_scaling.log 2>&1
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" >> $O/status.txt
cat $O/status.txt; cat $O/cfg4.log; cat $O/scan_scaling.log | tail -7; cat $O/bench_ref.json; tail -3 $O/bench_ref.err
